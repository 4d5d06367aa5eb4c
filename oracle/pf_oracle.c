/* TEST INFRASTRUCTURE (oracle). Not part of the product path. See pf_oracle.h for scope and pinning.
 *
 * Every function cites the reference file:line (relative to /root/reference/pathfinder) it restates.
 * Build: gcc -O2 -std=c11 -msse4.1 -ffp-contract=off (no FMA contraction: the x86 reference build has none,
 * CMakeLists.txt:58-62), default rounding mode, denormals kept.
 */
#include "pf_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define TILE 16
#define MAX_PAGES 64
#define MAX_BATCHES 64
#define CURVE_IS_QUADRATIC 0x80000000u /* core/data/data.h:19 */
#define CURVE_IS_CUBIC 0x40000000u     /* core/data/data.h:20 */
#define FLATTENING_TOLERANCE 1.0f      /* core/d3d9/tiler.cpp:15 */
#define FLOAT_EPSILON 0.0001f          /* common/math/basic.h:11 */
/* The reference recursion (tiler.cpp:284-315) and tile walk (tiler.cpp:191-278) are unbounded. Both this
 * oracle and the CUDA path bound them identically so that a hostile input cannot hang the GPU:
 * a curve still not flat at depth 24 is emitted as its baseline; a tile walk stops after 65536 steps. */
#define MAX_FLATTEN_DEPTH 24
#define MAX_DDA_STEPS 65536

typedef struct {
    float x, y;
} v2;

typedef struct {
    pfo_batch_desc desc;
    pfo_backdrop_info *backdrops;
    pfo_propagate_metadata *meta;
    pfo_dice_metadata *dice;
    pfo_tile_path_info *tpi;
    /* outputs */
    pfo_line *lines;
    size_t n_lines, cap_lines;
    pfo_line *clipped; /* lines that survive the view-box clip, clipped (what the tile walk consumes) */
    size_t n_clipped, cap_clipped;
    pfo_fill *fills;
    size_t n_fills, cap_fills;
    pfo_tile *tiles;      /* tile_count */
    int32_t *col_backdrop; /* column_count: crossings above the tile rect (object_builder.cpp:107-110) */
    int32_t *col_backdrop0; /* the same as bin left them (propagate advances col_backdrop down the columns) */
    int32_t *z11;          /* fb tiles */
    uint32_t *z9;
    uint32_t *list_offsets; /* fb tiles + 1 */
    uint32_t *list_tiles;
    uint32_t n_listed, n_listed_culled, max_list;
    uint32_t first_alpha, n_alpha;
    uint32_t *fill_offsets; /* tile_count + 1, into the sorted fills */
    int used;
} batch_t;

typedef struct {
    int w, h;
    uint8_t *px;
} page_t;

struct pfo_frame {
    int fb_w, fb_h, fb_tw, fb_th;
    int org_tx, org_ty; /* scene tile that maps to framebuffer tile (0, 0): horizontal strips of a large canvas */
    float view_box[4]; /* left top right bottom */
    uint8_t *lut;
    int lut_w, lut_h;
    float *points[2];
    uint32_t n_points[2];
    uint32_t *indices[2];
    uint32_t n_segments[2];
    uint16_t *metadata;
    uint32_t metadata_rows;
    page_t pages[MAX_PAGES];
    batch_t batches[MAX_BATCHES];
    int n_batches;
    uint8_t *masks; /* 256 B per alpha tile id */
    size_t n_masks, cap_masks;
    uint8_t *dest; /* RGBA8 */
    /* Pixel model: 0 = the GPU-driven shaders (fill.comp + tile.comp: RGBA8 mask with backdrop and fill rule applied, one
     * float accumulation per pixel, quantised once). 1 = the HYBRID raster shaders on the same geometry (shaders/d3d9/
     * fill.frag:24-45: raw signed areas blended additively into an RGBA16F mask, d3d9/renderer.cpp:152-158,220-222;
     * tile.frag sampleMask adds the backdrop and applies the fill rule at composite time; every tile primitive is a draw
     * into the RGBA8 target through the fixed-function src-over blender, gpu/base.h:83-93: quantised after every layer).
     * Model 1 exists to MEASURE the distance between the two variants the north star allows; clip combine passes
     * (tile_clip_combine.frag) are not modelled: scenes without clip paths only. */
    int pixel_model;
    float *masks16; /* model 1: raw area sums, every value representable as a half */
};

/* ------------------------------------------------------------------------------------------ frame plumbing */

pfo_frame *pfo_frame_create(int fb_width, int fb_height, const float view_box[4], const uint8_t *area_lut_rgba,
                            int lut_w, int lut_h) {
    pfo_frame *f = (pfo_frame *)calloc(1, sizeof(pfo_frame));
    f->fb_w = fb_width;
    f->fb_h = fb_height;
    /* core/d3d11/renderer.cpp:450-453 */
    f->fb_tw = (fb_width + TILE - 1) / TILE;
    f->fb_th = (fb_height + TILE - 1) / TILE;
    memcpy(f->view_box, view_box, sizeof(float) * 4);
    f->lut_w = lut_w;
    f->lut_h = lut_h;
    if (area_lut_rgba) {
        f->lut = (uint8_t *)malloc((size_t)lut_w * lut_h * 4);
        memcpy(f->lut, area_lut_rgba, (size_t)lut_w * lut_h * 4);
    }
    f->dest = (uint8_t *)calloc((size_t)fb_width * fb_height, 4);
    return f;
}

static void batch_free(batch_t *b) {
    free(b->backdrops);
    free(b->meta);
    free(b->dice);
    free(b->tpi);
    free(b->lines);
    free(b->clipped);
    free(b->fills);
    free(b->tiles);
    free(b->col_backdrop);
    free(b->col_backdrop0);
    free(b->z11);
    free(b->z9);
    free(b->list_offsets);
    free(b->list_tiles);
    free(b->fill_offsets);
    memset(b, 0, sizeof(*b));
}

void pfo_frame_reset(pfo_frame *f) {
    for (int i = 0; i < MAX_BATCHES; i++) batch_free(&f->batches[i]);
    f->n_batches = 0;
    f->n_masks = 0;
}

void pfo_frame_destroy(pfo_frame *f) {
    if (!f) return;
    pfo_frame_reset(f);
    for (int i = 0; i < 2; i++) {
        free(f->points[i]);
        free(f->indices[i]);
    }
    for (int i = 0; i < MAX_PAGES; i++) free(f->pages[i].px);
    free(f->metadata);
    free(f->lut);
    free(f->masks);
    free(f->masks16);
    free(f->dest);
    free(f);
}

void pfo_frame_set_segments(pfo_frame *f, int which, const float *points, uint32_t n_points,
                            const uint32_t *indices, uint32_t n_segments) {
    free(f->points[which]);
    free(f->indices[which]);
    f->points[which] = (float *)malloc(sizeof(float) * 2 * (n_points ? n_points : 1));
    f->indices[which] = (uint32_t *)malloc(sizeof(uint32_t) * 2 * (n_segments ? n_segments : 1));
    if (n_points) memcpy(f->points[which], points, sizeof(float) * 2 * n_points);
    if (n_segments) memcpy(f->indices[which], indices, sizeof(uint32_t) * 2 * n_segments);
    f->n_points[which] = n_points;
    f->n_segments[which] = n_segments;
}

void pfo_frame_set_metadata(pfo_frame *f, const uint16_t *half_texels, uint32_t n_rows) {
    free(f->metadata);
    size_t n = (size_t)n_rows * 1280 * 4;
    f->metadata = (uint16_t *)malloc(n * 2 + 2);
    memcpy(f->metadata, half_texels, n * 2);
    f->metadata_rows = n_rows;
}

void pfo_frame_set_page(pfo_frame *f, uint32_t page, int w, int h, const uint8_t *rgba) {
    if (page >= MAX_PAGES) return;
    free(f->pages[page].px);
    f->pages[page].w = w;
    f->pages[page].h = h;
    f->pages[page].px = (uint8_t *)calloc((size_t)w * h, 4);
    if (rgba) memcpy(f->pages[page].px, rgba, (size_t)w * h * 4);
}

/* ------------------------------------------------------------------------------------------ dice (flatten) */

static void push_line(batch_t *b, v2 from, v2 to, uint32_t path) {
    if (b->n_lines == b->cap_lines) {
        b->cap_lines = b->cap_lines ? b->cap_lines * 2 : 4096;
        b->lines = (pfo_line *)realloc(b->lines, b->cap_lines * sizeof(pfo_line));
    }
    pfo_line *l = &b->lines[b->n_lines++];
    l->from_x = from.x;
    l->from_y = from.y;
    l->to_x = to.x;
    l->to_y = to.y;
    l->path_index = path;
}

/* core/data/segment.cpp:12-20 (Segment::is_flat_cubic); F32x4 ops are lane-wise single precision. */
static int is_flat_cubic(v2 p0, v2 c0, v2 c1, v2 p3) {
    /* uv = 3 * ctrl - baseline - baseline - baseline.zwxy */
    float u0 = 3.0f * c0.x - p0.x - p0.x - p3.x;
    float u1 = 3.0f * c0.y - p0.y - p0.y - p3.y;
    float u2 = 3.0f * c1.x - p3.x - p3.x - p0.x;
    float u3 = 3.0f * c1.y - p3.y - p3.y - p0.y;
    u0 = u0 * u0;
    u1 = u1 * u1;
    u2 = u2 * u2;
    u3 = u3 * u3;
    /* _mm_max_ps(a, b) = a > b ? a : b (common/f32x4.h:52-54) */
    float m0 = u0 > u2 ? u0 : u2;
    float m1 = u1 > u3 ? u1 : u3;
    return m0 + m1 <= FLATTENING_TOLERANCE;
}

/* core/data/segment.cpp:22-37 (Segment::is_flat_quadratic) */
static int is_flat_quadratic(v2 p0, v2 p1, v2 p2) {
    float mx = (p0.x + p2.x) * 0.5f, my = (p0.y + p2.y) * 0.5f;
    float dx = p1.x - mx, dy = p1.y - my;
    return dx * dx + dy * dy <= FLATTENING_TOLERANCE * 0.25f;
}

static inline v2 lerp_half(v2 a, v2 b) { /* a + t * (b - a) with t = 0.5, segment.cpp:71-81 */
    v2 r = {a.x + 0.5f * (b.x - a.x), a.y + 0.5f * (b.y - a.y)};
    return r;
}

/* core/d3d9/tiler.cpp:284-315 (process_segment), cubic arm; split = segment.cpp:43-108 with t = 0.5. */
static void flatten_cubic(batch_t *b, v2 p0, v2 p1, v2 p2, v2 p3, uint32_t path, int depth) {
    if (depth >= MAX_FLATTEN_DEPTH || is_flat_cubic(p0, p1, p2, p3)) {
        push_line(b, p0, p3, path);
        return;
    }
    v2 p01 = lerp_half(p0, p1), p12 = lerp_half(p1, p2), p23 = lerp_half(p2, p3);
    v2 p012 = lerp_half(p01, p12), p123 = lerp_half(p12, p23);
    v2 p0123 = lerp_half(p012, p123);
    flatten_cubic(b, p0, p01, p012, p0123, path, depth + 1);
    flatten_cubic(b, p0123, p123, p23, p3, path, depth + 1);
}

/* tiler.cpp:291-303, quadratic arm; split = segment.cpp:110-135 (a = p0 + (p1 - p0) * t ...). */
static void flatten_quadratic(batch_t *b, v2 p0, v2 p1, v2 p2, uint32_t path, int depth) {
    if (depth >= MAX_FLATTEN_DEPTH || is_flat_quadratic(p0, p1, p2)) {
        push_line(b, p0, p2, path);
        return;
    }
    v2 a = {p0.x + (p1.x - p0.x) * 0.5f, p0.y + (p1.y - p0.y) * 0.5f};
    v2 bb = {p1.x + (p2.x - p1.x) * 0.5f, p1.y + (p2.y - p1.y) * 0.5f};
    v2 c = {a.x + (bb.x - a.x) * 0.5f, a.y + (bb.y - a.y) * 0.5f};
    flatten_quadratic(b, p0, a, c, path, depth + 1);
    flatten_quadratic(b, c, bb, p2, path, depth + 1);
}

static int finite2(v2 p) { return isfinite(p.x) && isfinite(p.y); }

/* The dice stage: walk the batch's segments (core/d3d11/gpu_data.cpp:79-115 layout) the way the hybrid
 * tiler walks a contour (tiler.cpp:345-367 + SegmentsIter, core/data/contour.cpp:110-173). */
static void dice_batch(pfo_frame *f, batch_t *b) {
    int which = b->desc.path_source;
    const float *pts = f->points[which];
    const uint32_t *idx = f->indices[which];
    uint32_t n_pts = f->n_points[which], n_seg = f->n_segments[which];
    const float *t = b->desc.transform;
    int identity = t[0] == 1.0f && t[1] == 0.0f && t[2] == 0.0f && t[3] == 1.0f && t[4] == 0.0f && t[5] == 0.0f;

    for (uint32_t p = 0; p < b->desc.path_count; p++) {
        uint32_t first_global = b->dice[p].first_global_segment_index;
        uint32_t first_batch = b->dice[p].first_batch_segment_index;
        uint32_t end_batch = p + 1 < b->desc.path_count ? b->dice[p + 1].first_batch_segment_index
                                                       : b->desc.segment_count;
        /* The contour start is needed for the closing-line test; track it while walking. */
        uint32_t contour_first_point = ~0u;
        for (uint32_t s = first_batch; s < end_batch; s++) {
            uint32_t g = first_global + (s - first_batch);
            if (g >= n_seg) break;
            uint32_t fp = idx[g * 2], flag = idx[g * 2 + 1];
            uint32_t next_fp = g + 1 < n_seg ? idx[(g + 1) * 2] : n_pts;
            if (contour_first_point == ~0u) contour_first_point = fp;
            v2 q[4];
            uint32_t npt = (flag & CURVE_IS_CUBIC) ? 4 : (flag & CURVE_IS_QUADRATIC) ? 3 : 2;
            if (fp + npt > n_pts) break;
            for (uint32_t k = 0; k < npt; k++) {
                float x = pts[(fp + k) * 2], y = pts[(fp + k) * 2 + 1];
                if (!identity) { /* common/math/transform2.h:89-91, mat2.h:84-86 */
                    float tx = t[0] * x + t[2] * y + t[4];
                    float ty = t[1] * x + t[3] * y + t[5];
                    x = tx;
                    y = ty;
                }
                q[k].x = x;
                q[k].y = y;
            }
            /* SegmentsD3D11::add_path appends points[0] after each contour (gpu_data.cpp:109), so a line whose
             * successor starts two points later is the contour's closing line. The hybrid tiler only emits it
             * when the contour is not already closed (contour.cpp:157-166: approx_eq(front, back, 1e-4)). */
            int closing = npt == 2 && next_fp == fp + 2;
            if (closing) {
                contour_first_point = ~0u;
                float dx = q[1].x - q[0].x, dy = q[1].y - q[0].y; /* (front - back).length() <= eps */
                if (sqrtf(dx * dx + dy * dy) <= FLOAT_EPSILON) continue;
            }
            int ok = 1;
            for (uint32_t k = 0; k < npt; k++) ok &= finite2(q[k]);
            if (!ok) continue; /* core/data/line_segment.cpp:97-104 rejects non-finite lines */
            if (npt == 2) {
                push_line(b, q[0], q[1], p);
            } else if (npt == 3) {
                flatten_quadratic(b, q[0], q[1], q[2], p, 0);
            } else {
                flatten_cubic(b, q[0], q[1], q[2], q[3], p, 0);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ bin (tile) */

static void push_fill(batch_t *b, uint32_t tile_index, uint16_t fx, uint16_t fy, uint16_t tx, uint16_t ty) {
    if (b->n_fills == b->cap_fills) {
        b->cap_fills = b->cap_fills ? b->cap_fills * 2 : 8192;
        b->fills = (pfo_fill *)realloc(b->fills, b->cap_fills * sizeof(pfo_fill));
    }
    pfo_fill *q = &b->fills[b->n_fills++];
    q->tile_index = tile_index;
    q->from_x = fx;
    q->from_y = fy;
    q->to_x = tx;
    q->to_y = ty;
}

/* core/d3d9/object_builder.cpp:19-64 (ObjectBuilder::add_fill) */
static void add_fill(batch_t *b, const pfo_propagate_metadata *m, float fx, float fy, float tx, float ty, int tcx,
                     int tcy) {
    /* RectI::contains_point(Vec2I), common/math/rect.h:195-198 */
    if (!(m->rect[0] <= tcx && tcx <= m->rect[2] - 1 && m->rect[1] <= tcy && tcy <= m->rect[3] - 1)) return;
    float ulx = (float)tcx * (float)TILE, uly = (float)tcy * (float)TILE;
    float s[4] = {(fx - ulx) * 256.0f, (fy - uly) * 256.0f, (tx - ulx) * 256.0f, (ty - uly) * 256.0f};
    uint16_t u[4];
    for (int i = 0; i < 4; i++) {
        float v = s[i];
        v = v > 0.0f ? v : 0.0f;       /* _mm_max_ps(v, 0) */
        v = v < 4095.0f ? v : 4095.0f; /* _mm_min_ps(v, 16 * 256 - 1) */
        v = rintf(v);                  /* _mm_round_ps(TO_NEAREST_INT): ties to even (common/f32x4.h:64-66) */
        u[i] = (uint16_t)v;
    }
    if (u[0] == u[2]) return; /* cull degenerate fills */
    uint32_t tile_index = m->tile_offset + (uint32_t)(tcx - m->rect[0]) +
                          (uint32_t)(m->rect[2] - m->rect[0]) * (uint32_t)(tcy - m->rect[1]);
    push_fill(b, tile_index, u[0], u[1], u[2], u[3]);
}

/* core/d3d9/object_builder.cpp:95-114 (ObjectBuilder::adjust_alpha_tile_backdrop) */
static void adjust_backdrop(batch_t *b, const pfo_propagate_metadata *m, int tcx, int tcy, int delta) {
    int ox = tcx - m->rect[0], oy = tcy - m->rect[1];
    int w = m->rect[2] - m->rect[0], h = m->rect[3] - m->rect[1];
    if (ox < 0 || ox >= w || oy >= h) return;
    if (oy < 0) {
        b->col_backdrop[m->backdrop_offset + (uint32_t)ox] += delta;
        return;
    }
    pfo_tile *t = &b->tiles[m->tile_offset + (uint32_t)ox + (uint32_t)w * (uint32_t)oy];
    t->backdrop_delta = (int8_t)(t->backdrop_delta + delta);
}

static inline float clampf01(float t) { return t < 0.0f ? 0.0f : (t > 1.0f ? 1.0f : t); }
/* common/math/basic.h:50-53 (scalar lerp clamps t) */
static inline float lerp_clamped(float a, float b, float t) {
    t = clampf01(t);
    return a + (b - a) * t;
}

/* core/d3d9/tiler.cpp:46-62 */
static unsigned outcode(float x, float y, const float r[4]) {
    unsigned c = 0;
    if (x < r[0]) c |= 1; /* LEFT */
    else if (x > r[2]) c |= 2; /* RIGHT */
    if (y < r[1]) c |= 4; /* TOP */
    else if (y > r[3]) c |= 8; /* BOTTOM */
    return c;
}

/* core/d3d9/tiler.cpp:65-125 (clip_line_segment_to_rect, Cohen-Sutherland) */
static int clip_line(float *l, const float r[4]) {
    unsigned of = outcode(l[0], l[1], r), ot = outcode(l[2], l[3], r);
    for (int guard = 0; guard < 16; guard++) {
        if (of == 0 && ot == 0) return 1;
        if ((of & ot) != 0) return 0;
        int clip_from = of > ot;
        unsigned oc = clip_from ? of : ot;
        float px = clip_from ? l[0] : l[2], py = clip_from ? l[1] : l[3];
        if (oc & 1) {
            py = lerp_clamped(l[1], l[3], (r[0] - l[0]) / (l[2] - l[0]));
            px = r[0];
        } else if (oc & 2) {
            py = lerp_clamped(l[1], l[3], (r[2] - l[0]) / (l[2] - l[0]));
            px = r[2];
        } else if (oc & 4) {
            px = lerp_clamped(l[0], l[2], (r[1] - l[1]) / (l[3] - l[1]));
            py = r[1];
        } else if (oc & 8) {
            px = lerp_clamped(l[0], l[2], (r[3] - l[1]) / (l[3] - l[1]));
            py = r[3];
        }
        if (clip_from) {
            l[0] = px;
            l[1] = py;
            of = outcode(px, py, r);
        } else {
            l[2] = px;
            l[3] = py;
            ot = outcode(px, py, r);
        }
    }
    return 0;
}

/* core/d3d9/tiler.cpp:131-279 (process_line_segment) */
static void process_line(pfo_frame *f, batch_t *b, const pfo_line *ln) {
    const pfo_propagate_metadata *m = &b->meta[ln->path_index];
    float l[4] = {ln->from_x, ln->from_y, ln->to_x, ln->to_y};
    if (!(isfinite(l[0]) && isfinite(l[1]) && isfinite(l[2]) && isfinite(l[3]))) return;
    /* view box with an open top (tiler.cpp:141-152) */
    float box[4] = {f->view_box[0], -INFINITY, f->view_box[2], f->view_box[3]};
    if (!clip_line(l, box)) return;
    if (b->n_clipped == b->cap_clipped) {
        b->cap_clipped = b->cap_clipped ? b->cap_clipped * 2 : 4096;
        b->clipped = (pfo_line *)realloc(b->clipped, b->cap_clipped * sizeof(pfo_line));
    }
    {
        pfo_line *cl = &b->clipped[b->n_clipped++];
        cl->from_x = l[0];
        cl->from_y = l[1];
        cl->to_x = l[2];
        cl->to_y = l[3];
        cl->path_index = ln->path_index;
    }

    const float ts = (float)TILE;
    float tlx = l[0] * (1.0f / TILE), tly = l[1] * (1.0f / TILE);
    float ttx = l[2] * (1.0f / TILE), tty = l[3] * (1.0f / TILE);
    int from_tx = (int)floorf(tlx), from_ty = (int)floorf(tly);
    int to_tx = (int)floorf(ttx), to_ty = (int)floorf(tty);
    float vx = l[2] - l[0], vy = l[3] - l[1];
    int step_x = vx < 0 ? -1 : 1, step_y = vy < 0 ? -1 : 1;
    float fcx = ((float)from_tx + (vx >= 0 ? 1.0f : 0.0f)) * ts;
    float fcy = ((float)from_ty + (vy >= 0 ? 1.0f : 0.0f)) * ts;
    float t_max_x = (fcx - l[0]) / vx, t_max_y = (fcy - l[1]) / vy;
    float t_delta_x = fabsf(ts / vx), t_delta_y = fabsf(ts / vy);
    float cur_x = l[0], cur_y = l[1];
    int tcx = from_tx, tcy = from_ty;
    int last_dir = 0; /* 0 none, 1 X, 2 Y */

    for (int iter = 0; iter < MAX_DDA_STEPS; iter++) {
        int next_dir;
        if (t_max_x < t_max_y) next_dir = 1;
        else if (t_max_x > t_max_y) next_dir = 2;
        else next_dir = step_x > 0 ? 1 : 2;
        float next_t = next_dir == 1 ? t_max_x : t_max_y;
        next_t = next_t < 1.0f ? next_t : 1.0f; /* std::min(next_t, 1.0f) */
        if (tcx == to_tx && tcy == to_ty) next_dir = 0;
        /* LineSegmentF::sample: from + vector * t (core/data/line_segment.h:58-60) */
        float nx = l[0] + vx * next_t, ny = l[1] + vy * next_t;
        add_fill(b, m, cur_x, cur_y, nx, ny, tcx, tcy);
        if (step_y < 0 && next_dir == 2) {
            add_fill(b, m, nx, ny, (float)tcx * ts, (float)tcy * ts, tcx, tcy);
        } else if (step_y > 0 && last_dir == 2) {
            add_fill(b, m, (float)tcx * ts, (float)tcy * ts, cur_x, cur_y, tcx, tcy);
        }
        if (step_x < 0 && last_dir == 1) adjust_backdrop(b, m, tcx, tcy, 1);
        else if (step_x > 0 && next_dir == 1) adjust_backdrop(b, m, tcx, tcy, -1);
        if (next_dir == 1) {
            t_max_x += t_delta_x;
            tcx += step_x;
        } else if (next_dir == 2) {
            t_max_y += t_delta_y;
            tcy += step_y;
        } else {
            break;
        }
        cur_x = nx;
        cur_y = ny;
        last_dir = next_dir;
    }
}

static int fill_cmp(const void *a, const void *b) {
    const pfo_fill *x = (const pfo_fill *)a, *y = (const pfo_fill *)b;
    if (x->tile_index != y->tile_index) return x->tile_index < y->tile_index ? -1 : 1;
    if (x->from_x != y->from_x) return x->from_x < y->from_x ? -1 : 1;
    if (x->from_y != y->from_y) return x->from_y < y->from_y ? -1 : 1;
    if (x->to_x != y->to_x) return x->to_x < y->to_x ? -1 : 1;
    if (x->to_y != y->to_y) return x->to_y < y->to_y ? -1 : 1;
    return 0;
}

/* ------------------------------------------------------------------------------------------ propagate */

static batch_t *find_batch(pfo_frame *f, uint32_t batch_id) {
    for (int i = 0; i < f->n_batches; i++)
        if (f->batches[i].used && f->batches[i].desc.batch_id == batch_id) return &f->batches[i];
    return NULL;
}

/* core/d3d9/tiler.cpp:369-439 (Tiler::prepare_tiles) in the data model of shaders/d3d11/propagate.comp:95-216. */
static void propagate_batch(pfo_frame *f, batch_t *b) {
    batch_t *cb = b->desc.clip_batch_id >= 0 ? find_batch(f, (uint32_t)b->desc.clip_batch_id) : NULL;
    size_t fbt = (size_t)f->fb_tw * f->fb_th;
    b->z11 = (int32_t *)calloc(fbt ? fbt : 1, sizeof(int32_t));
    b->z9 = (uint32_t *)calloc(fbt ? fbt : 1, sizeof(uint32_t));
    uint32_t *list_count = (uint32_t *)calloc(fbt + 1, sizeof(uint32_t));
    b->first_alpha = (uint32_t)f->n_masks;
    uint32_t next_alpha = b->first_alpha;
    free(b->col_backdrop0);
    b->col_backdrop0 = (int32_t *)malloc(sizeof(int32_t) * (b->desc.column_count ? b->desc.column_count : 1));
    memcpy(b->col_backdrop0, b->col_backdrop, sizeof(int32_t) * b->desc.column_count);

    /* Alpha tile ids are allocated in dense tile order (deterministic stand-in for propagate.comp:178-183's
     * atomicAdd); backdrops are a prefix sum down each column. Walk row-major like tiler.cpp:381. */
    for (uint32_t p = 0; p < b->desc.path_count; p++) {
        const pfo_propagate_metadata *m = &b->meta[p];
        int w = m->rect[2] - m->rect[0], h = m->rect[3] - m->rect[1];
        if (w <= 0 || h <= 0) continue;
        const pfo_propagate_metadata *cm = NULL;
        if (cb && (int32_t)m->clip_path_index >= 0 && m->clip_path_index < cb->desc.path_count)
            cm = &cb->meta[m->clip_path_index];
        int even_odd = (b->tpi[p].ctrl & 0x2) != 0;
        for (int ty = 0; ty < h; ty++) {
            for (int tx = 0; tx < w; tx++) {
                uint32_t ti = m->tile_offset + (uint32_t)tx + (uint32_t)w * (uint32_t)ty;
                pfo_tile *t = &b->tiles[ti];
                int32_t *col = &b->col_backdrop[m->backdrop_offset + (uint32_t)tx];
                int backdrop = (int8_t)*col; /* int8_t(backdrops[column]), tiler.cpp:394 */
                int backdrop9 = backdrop;
                int have_mask = t->fill_count > 0;
                int need_new = have_mask;
                int alpha = -1, clip_alpha = -1;
                int has_alpha9 = have_mask;
                if ((int32_t)m->clip_path_index >= 0) { /* a clip is applied */
                    int gx = tx + m->rect[0], gy = ty + m->rect[1];
                    int inside = cm && gx >= cm->rect[0] && gx < cm->rect[2] && gy >= cm->rect[1] && gy < cm->rect[3];
                    if (inside) {
                        const pfo_tile *ct = &cb->tiles[cm->tile_offset + (uint32_t)(gx - cm->rect[0]) +
                                                        (uint32_t)(cm->rect[2] - cm->rect[0]) *
                                                            (uint32_t)(gy - cm->rect[1])];
                        if (ct->alpha_tile_id >= 0) {
                            if (have_mask) { /* tiler.cpp:403-414 / propagate.comp:144-147 */
                                clip_alpha = ct->alpha_tile_id;
                                need_new = 1;
                                backdrop9 = 0;
                            } else if (backdrop != 0) { /* tiler.cpp:415-420 / propagate.comp:149-154 */
                                alpha = ct->alpha_tile_id;
                                need_new = 0;
                                has_alpha9 = 1;
                                backdrop9 = ct->backdrop_d3d9;
                            } else {
                                need_new = 0;
                            }
                        } else if (ct->backdrop == 0) { /* blank clip tile: tiler.cpp:421-425 / propagate.comp:163-169 */
                            backdrop = 0;
                            backdrop9 = 0;
                            need_new = 0;
                            has_alpha9 = 0;
                        }
                    } else { /* outside the clip rect: tiler.cpp:426-430 / propagate.comp:171-175 */
                        backdrop = 0;
                        backdrop9 = 0;
                        need_new = 0;
                        has_alpha9 = 0;
                    }
                }
                if (need_new) alpha = (int)next_alpha++;
                t->alpha_tile_id = alpha;
                t->clip_alpha_tile_id = need_new ? clip_alpha : -1;
                t->backdrop = (int8_t)backdrop;
                t->backdrop_d3d9 = (int8_t)backdrop9;
                (void)has_alpha9;

                int gx = tx + m->rect[0] - f->org_tx, gy = ty + m->rect[1] - f->org_ty;
                int in_fb = gx >= 0 && gx < f->fb_tw && gy >= 0 && gy < f->fb_th;
                size_t map = (size_t)gy * f->fb_tw + gx;
                /* z: propagate.comp:190-206 */
                int z_write = m->z_write != 0;
                if ((int8_t)backdrop != 0 && even_odd && (abs((int8_t)backdrop) % 2) == 0) z_write = 0;
                if (in_fb && z_write && (int8_t)backdrop != 0 && alpha < 0) {
                    if ((int32_t)ti > b->z11[map]) b->z11[map] = (int32_t)ti;
                }
                /* hybrid z: core/d3d9/scene_builder.cpp:57-78 (max draw path id over occluding solid tiles) */
                if (in_fb && m->z_write && (int8_t)backdrop9 != 0 && !(alpha >= 0)) {
                    uint32_t gid = b->dice[p].global_path_id;
                    if (gid > b->z9[map]) b->z9[map] = gid;
                }
                t->listed = (uint8_t)(((int8_t)backdrop != 0 || alpha >= 0) && in_fb);
                if (t->listed) {
                    list_count[map]++;
                    b->n_listed++;
                }
                *col += t->backdrop_delta; /* tiler.cpp:437 */
            }
        }
    }
    b->n_alpha = next_alpha - b->first_alpha;
    /* grow the mask store */
    if (next_alpha > f->cap_masks) {
        f->cap_masks = next_alpha * 2 + 64;
        f->masks = (uint8_t *)realloc(f->masks, f->cap_masks * 256);
        f->masks16 = (float *)realloc(f->masks16, f->cap_masks * 256 * sizeof(float));
    }
    f->n_masks = next_alpha;

    /* sort.comp:49-83: per framebuffer tile, ascending dense tile index, dropping entries below z. */
    b->list_offsets = (uint32_t *)calloc(fbt + 1, sizeof(uint32_t));
    b->list_tiles = (uint32_t *)malloc(sizeof(uint32_t) * (b->n_listed ? b->n_listed : 1));
    uint32_t *cursor = (uint32_t *)calloc(fbt + 1, sizeof(uint32_t));
    /* first pass: count survivors */
    memset(list_count, 0, sizeof(uint32_t) * (fbt + 1));
    for (uint32_t p = 0; p < b->desc.path_count; p++) {
        const pfo_propagate_metadata *m = &b->meta[p];
        int w = m->rect[2] - m->rect[0], h = m->rect[3] - m->rect[1];
        for (int ty = 0; ty < h; ty++)
            for (int tx = 0; tx < w; tx++) {
                uint32_t ti = m->tile_offset + (uint32_t)tx + (uint32_t)w * (uint32_t)ty;
                if (!b->tiles[ti].listed) continue;
                size_t map = (size_t)(ty + m->rect[1] - f->org_ty) * f->fb_tw + (tx + m->rect[0] - f->org_tx);
                if ((int32_t)ti >= b->z11[map]) list_count[map]++;
            }
    }
    uint32_t acc = 0;
    for (size_t i = 0; i < fbt; i++) {
        b->list_offsets[i] = acc;
        acc += list_count[i];
        if (list_count[i] > b->max_list) b->max_list = list_count[i];
    }
    b->list_offsets[fbt] = acc;
    b->n_listed_culled = acc;
    /* dense tile order is already ascending within a path and paths ascend, so appending keeps lists sorted */
    for (uint32_t p = 0; p < b->desc.path_count; p++) {
        const pfo_propagate_metadata *m = &b->meta[p];
        int w = m->rect[2] - m->rect[0], h = m->rect[3] - m->rect[1];
        for (int ty = 0; ty < h; ty++)
            for (int tx = 0; tx < w; tx++) {
                uint32_t ti = m->tile_offset + (uint32_t)tx + (uint32_t)w * (uint32_t)ty;
                if (!b->tiles[ti].listed) continue;
                size_t map = (size_t)(ty + m->rect[1] - f->org_ty) * f->fb_tw + (tx + m->rect[0] - f->org_tx);
                if ((int32_t)ti >= b->z11[map]) b->list_tiles[b->list_offsets[map] + cursor[map]++] = ti;
            }
    }
    free(cursor);
    free(list_count);
}

/* ------------------------------------------------------------------------------------------ fill (masks) */

/* bilinear, clamp-to-edge, unnormalised texel-space sample of an RGBA8 image; returns [0,1] floats */
static void sample_rgba8(const uint8_t *px, int w, int h, float u, float v, int repeat_u, int repeat_v, int nearest,
                         float out[4]) {
    /* u, v are normalised texture coordinates */
    float x = u * (float)w, y = v * (float)h;
    if (nearest) {
        int ix = (int)floorf(x), iy = (int)floorf(y);
        if (repeat_u) ix = ((ix % w) + w) % w; else ix = ix < 0 ? 0 : (ix >= w ? w - 1 : ix);
        if (repeat_v) iy = ((iy % h) + h) % h; else iy = iy < 0 ? 0 : (iy >= h ? h - 1 : iy);
        const uint8_t *t = px + ((size_t)iy * w + ix) * 4;
        for (int c = 0; c < 4; c++) out[c] = (float)t[c] * (1.0f / 255.0f);
        return;
    }
    x -= 0.5f;
    y -= 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float ax = x - fx0, ay = y - fy0;
    int x0 = (int)fx0, y0 = (int)fy0, x1 = x0 + 1, y1 = y0 + 1;
    if (repeat_u) {
        x0 = ((x0 % w) + w) % w;
        x1 = ((x1 % w) + w) % w;
    } else {
        x0 = x0 < 0 ? 0 : (x0 >= w ? w - 1 : x0);
        x1 = x1 < 0 ? 0 : (x1 >= w ? w - 1 : x1);
    }
    if (repeat_v) {
        y0 = ((y0 % h) + h) % h;
        y1 = ((y1 % h) + h) % h;
    } else {
        y0 = y0 < 0 ? 0 : (y0 >= h ? h - 1 : y0);
        y1 = y1 < 0 ? 0 : (y1 >= h ? h - 1 : y1);
    }
    const uint8_t *t00 = px + ((size_t)y0 * w + x0) * 4, *t10 = px + ((size_t)y0 * w + x1) * 4;
    const uint8_t *t01 = px + ((size_t)y1 * w + x0) * 4, *t11 = px + ((size_t)y1 * w + x1) * 4;
    for (int c = 0; c < 4; c++) {
        float a = (float)t00[c] * (1.0f / 255.0f), bq = (float)t10[c] * (1.0f / 255.0f);
        float cq = (float)t01[c] * (1.0f / 255.0f), d = (float)t11[c] * (1.0f / 255.0f);
        float top = a + (bq - a) * ax, bot = cq + (d - cq) * ax;
        out[c] = top + (bot - top) * ay;
    }
}

static inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; } /* GLSL mix */
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* shaders/d3d11/fill.comp:51-71 (computeCoverage): 4 vertically adjacent pixels per call */
static void compute_coverage(const pfo_frame *f, float fx, float fy, float tx, float ty, float cov[4]) {
    float lx, ly, rx, ry;
    if (fx < tx) {
        lx = fx; ly = fy; rx = tx; ry = ty;
    } else {
        lx = tx; ly = ty; rx = fx; ry = fy;
    }
    float wx = clampf(fx, -0.5f, 0.5f), wy = clampf(tx, -0.5f, 0.5f);
    float offset = mixf(wx, wy, 0.5f) - lx;
    float t = offset / (rx - lx);
    float y = mixf(ly, ry, t);
    float d = (ry - ly) / (rx - lx);
    float dX = wx - wy;
    float s[4];
    sample_rgba8(f->lut, f->lut_w, f->lut_h, (y + 8.0f) / 16.0f, fabsf(d * dX) / 16.0f, 0, 0, 0, s);
    for (int c = 0; c < 4; c++) cov[c] = s[c] * dX;
}

static inline float glsl_mod(float x, float y) { return x - y * floorf(x / y); }

/* float -> nearest half (round to nearest even) -> float: what storing into an RGBA16F render target does */
static float round_to_half(float v) {
    union { float f; uint32_t u; } x;
    x.f = v;
    const uint32_t sign = x.u & 0x80000000u;
    x.u &= 0x7fffffffu;
    if (x.u >= 0x7f800000u) return v;                 /* inf / nan */
    if (x.f >= 65520.0f) { x.u = 0x7f800000u | sign; return x.f; } /* overflows half */
    if (x.f < 6.103515625e-05f) {                     /* half subnormal: multiples of 2^-24 */
        const float q = rintf(x.f * 16777216.0f) / 16777216.0f;
        x.f = q;
        x.u |= sign;
        return x.f;
    }
    const uint32_t rem = x.u & 0x1fffu;               /* 13 dropped mantissa bits */
    x.u &= ~0x1fffu;
    if (rem > 0x1000u || (rem == 0x1000u && (x.u & 0x2000u))) x.u += 0x2000u;
    x.u |= sign;
    return x.f;
}

/* shaders/d3d11/fill.comp:109-154 (main) for every alpha tile this batch allocated */
static void fill_batch(pfo_frame *f, batch_t *b) {
    for (uint32_t ti = 0; ti < b->desc.tile_count; ti++) {
        const pfo_tile *t = &b->tiles[ti];
        if (t->alpha_tile_id < 0 || (uint32_t)t->alpha_tile_id < b->first_alpha) continue; /* reuses a clip slot */
        uint8_t *mask = f->masks + (size_t)t->alpha_tile_id * 256;
        const pfo_fill *fl = b->fills + b->fill_offsets[ti];
        uint32_t nf = b->fill_offsets[ti + 1] - b->fill_offsets[ti];
        /* tile ctrl comes from the path (bound.comp:72) */
        uint32_t path = 0;
        { /* find the path that owns this tile (bound.comp:41-53 does a binary search) */
            uint32_t lo = 0, hi = b->desc.path_count;
            while (lo + 1 < hi) {
                uint32_t mid = lo + (hi - lo) / 2;
                if (ti < b->meta[mid].tile_offset) hi = mid; else lo = mid;
            }
            path = lo;
        }
        int ctrl = b->tpi[path].ctrl;
        for (int lyq = 0; lyq < 4; lyq++) {
            for (int lxq = 0; lxq < 16; lxq++) {
                float fragx = (float)lxq + 0.5f, fragy = (float)(lyq * 4) + 0.5f;
                float cov[4] = {(float)t->backdrop, (float)t->backdrop, (float)t->backdrop, (float)t->backdrop};
                for (uint32_t k = 0; k < nf; k++) {
                    float c[4];
                    compute_coverage(f, (float)fl[k].from_x / 256.0f - fragx, (float)fl[k].from_y / 256.0f - fragy,
                                     (float)fl[k].to_x / 256.0f - fragx, (float)fl[k].to_y / 256.0f - fragy, c);
                    for (int q = 0; q < 4; q++) cov[q] += c[q];
                }
                if (f->pixel_model == 1) { /* additive blending into RGBA16F: the running sum is a half after every fill */
                    float raw[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                    for (uint32_t k = 0; k < nf; k++) {
                        float c[4];
                        compute_coverage(f, (float)fl[k].from_x / 256.0f - fragx, (float)fl[k].from_y / 256.0f - fragy,
                                         (float)fl[k].to_x / 256.0f - fragx, (float)fl[k].to_y / 256.0f - fragy, c);
                        for (int q = 0; q < 4; q++) raw[q] = round_to_half(raw[q] + c[q]);
                    }
                    for (int q = 0; q < 4; q++) f->masks16[(size_t)t->alpha_tile_id * 256 + (lyq * 4 + q) * 16 + lxq] = raw[q];
                }
                for (int q = 0; q < 4; q++) {
                    float cv = cov[q];
                    if (ctrl & 0x1) cv = clampf(fabsf(cv), 0.0f, 1.0f);
                    else cv = clampf(1.0f - fabsf(1.0f - glsl_mod(cv, 2.0f)), 0.0f, 1.0f);
                    int row = lyq * 4 + q;
                    if (t->clip_alpha_tile_id >= 0) {
                        float cl = (float)f->masks[(size_t)t->clip_alpha_tile_id * 256 + row * 16 + lxq] *
                                   (1.0f / 255.0f);
                        cv = cv < cl ? cv : cl;
                    }
                    mask[row * 16 + lxq] = (uint8_t)rintf(cv * 255.0f); /* rgba8 unorm store */
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------ batch driver */

static int prepare_batch(pfo_frame *f, const pfo_batch_desc *d, int with_masks) {
    if (f->n_batches >= MAX_BATCHES) return -1;
    int slot = f->n_batches++;
    batch_t *b = &f->batches[slot];
    memset(b, 0, sizeof(*b));
    b->used = 1;
    b->desc = *d;
    b->backdrops = (pfo_backdrop_info *)malloc(sizeof(pfo_backdrop_info) * (d->column_count ? d->column_count : 1));
    b->meta = (pfo_propagate_metadata *)malloc(sizeof(pfo_propagate_metadata) * (d->path_count ? d->path_count : 1));
    b->dice = (pfo_dice_metadata *)malloc(sizeof(pfo_dice_metadata) * (d->path_count ? d->path_count : 1));
    b->tpi = (pfo_tile_path_info *)malloc(sizeof(pfo_tile_path_info) * (d->path_count ? d->path_count : 1));
    memcpy(b->backdrops, d->backdrops, sizeof(pfo_backdrop_info) * d->column_count);
    memcpy(b->meta, d->propagate_metadata, sizeof(pfo_propagate_metadata) * d->path_count);
    memcpy(b->dice, d->dice_metadata, sizeof(pfo_dice_metadata) * d->path_count);
    memcpy(b->tpi, d->tile_path_info, sizeof(pfo_tile_path_info) * d->path_count);
    b->desc.backdrops = b->backdrops;
    b->desc.propagate_metadata = b->meta;
    b->desc.dice_metadata = b->dice;
    b->desc.tile_path_info = b->tpi;
    b->tiles = (pfo_tile *)calloc(d->tile_count ? d->tile_count : 1, sizeof(pfo_tile));
    b->col_backdrop = (int32_t *)calloc(d->column_count ? d->column_count : 1, sizeof(int32_t));
    for (uint32_t c = 0; c < d->column_count; c++) b->col_backdrop[c] = b->backdrops[c].initial_backdrop;

    dice_batch(f, b);
    for (size_t i = 0; i < b->n_lines; i++) process_line(f, b, &b->lines[i]);
    qsort(b->fills, b->n_fills, sizeof(pfo_fill), fill_cmp);
    b->fill_offsets = (uint32_t *)calloc((size_t)d->tile_count + 2, sizeof(uint32_t));
    for (size_t i = 0; i < b->n_fills; i++) b->fill_offsets[b->fills[i].tile_index + 1]++;
    for (uint32_t i = 0; i < d->tile_count; i++) {
        b->tiles[i].fill_count = (int32_t)b->fill_offsets[i + 1];
        b->fill_offsets[i + 1] += b->fill_offsets[i];
    }
    propagate_batch(f, b);
    if (with_masks && f->lut) fill_batch(f, b);
    return slot;
}

int pfo_frame_prepare_batch(pfo_frame *f, const pfo_batch_desc *desc) { return prepare_batch(f, desc, 1); }
int pfo_frame_prepare_batch_geometry_only(pfo_frame *f, const pfo_batch_desc *desc) {
    return prepare_batch(f, desc, 0);
}

void pfo_batch_counts(const pfo_frame *f, int slot, uint32_t c[8]) {
    const batch_t *b = &f->batches[slot];
    memset(c, 0, sizeof(uint32_t) * 8);
    c[0] = (uint32_t)b->n_lines;
    c[1] = (uint32_t)b->n_fills;
    c[2] = b->n_alpha;
    c[3] = b->first_alpha;
    c[4] = b->n_listed;
    c[5] = b->n_listed_culled;
    c[6] = b->max_list;
}

size_t pfo_batch_lines(const pfo_frame *f, int slot, pfo_line *out) {
    const batch_t *b = &f->batches[slot];
    if (out) memcpy(out, b->lines, b->n_lines * sizeof(pfo_line));
    return b->n_lines;
}

size_t pfo_batch_clipped_lines(const pfo_frame *f, int slot, pfo_line *out) {
    const batch_t *b = &f->batches[slot];
    if (out) memcpy(out, b->clipped, b->n_clipped * sizeof(pfo_line));
    return b->n_clipped;
}

size_t pfo_batch_fills(const pfo_frame *f, int slot, pfo_fill *out) {
    const batch_t *b = &f->batches[slot];
    if (out) memcpy(out, b->fills, b->n_fills * sizeof(pfo_fill));
    return b->n_fills;
}

size_t pfo_batch_tiles(const pfo_frame *f, int slot, pfo_tile *out) {
    const batch_t *b = &f->batches[slot];
    if (out) memcpy(out, b->tiles, (size_t)b->desc.tile_count * sizeof(pfo_tile));
    return b->desc.tile_count;
}

size_t pfo_batch_column_backdrops(const pfo_frame *f, int slot, int32_t *out) {
    if (slot < 0 || slot >= f->n_batches) return 0;
    const batch_t *b = &f->batches[slot];
    if (out && b->col_backdrop0) memcpy(out, b->col_backdrop0, sizeof(int32_t) * b->desc.column_count);
    return b->desc.column_count;
}

size_t pfo_batch_z(const pfo_frame *f, int slot, int32_t *z11, uint32_t *z9) {
    const batch_t *b = &f->batches[slot];
    size_t n = (size_t)f->fb_tw * f->fb_th;
    if (z11) memcpy(z11, b->z11, n * 4);
    if (z9) memcpy(z9, b->z9, n * 4);
    return n;
}

size_t pfo_batch_tile_lists(const pfo_frame *f, int slot, uint32_t *offsets, uint32_t *tiles) {
    const batch_t *b = &f->batches[slot];
    size_t n = (size_t)f->fb_tw * f->fb_th;
    if (offsets) memcpy(offsets, b->list_offsets, (n + 1) * 4);
    if (tiles) memcpy(tiles, b->list_tiles, (size_t)b->n_listed_culled * 4);
    return b->n_listed_culled;
}

int pfo_frame_mask(const pfo_frame *f, uint32_t id, uint8_t out[256]) {
    if (id >= f->n_masks) return -1;
    memcpy(out, f->masks + (size_t)id * 256, 256);
    return 0;
}

/* ------------------------------------------------------------------------------------------ tile (composite) */

static float half_to_float(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000) << 16, exp = (h >> 10) & 0x1f, man = h & 0x3ff, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {
            int e = -1;
            do {
                e++;
                man <<= 1;
            } while (!(man & 0x400));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112) << 23) | (man << 13);
    float fl;
    memcpy(&fl, &bits, 4);
    return fl;
}

typedef struct {
    float v[4];
} vec4;

/* tile.comp:690-692,707-717 (fetchUnscaled at texel centres == exact texel) */
static vec4 metadata_texel(const pfo_frame *f, int color_entry, int entry) {
    vec4 r = {{0, 0, 0, 0}};
    int x = color_entry % 128 * 10 + entry, y = color_entry / 128;
    if ((uint32_t)y >= f->metadata_rows) return r;
    const uint16_t *t = f->metadata + ((size_t)y * 1280 + x) * 4;
    for (int c = 0; c < 4; c++) r.v[c] = half_to_float(t[c]);
    return r;
}

typedef struct {
    const uint8_t *px;
    int w, h, repeat_u, repeat_v, nearest;
} sampler_t;

static vec4 tex(const sampler_t *s, float u, float v) {
    vec4 r = {{0, 0, 0, 0}};
    if (!s->px) return r;
    sample_rgba8(s->px, s->w, s->h, u, v, s->repeat_u, s->repeat_v, s->nearest, r.v);
    return r;
}

/* tile.comp:319-347 (filterRadialGradient) */
static vec4 filter_radial(const sampler_t *s, float cu, float cv, vec4 p0, vec4 p1) {
    float lfx = p0.v[0], lfy = p0.v[1], lvx = p0.v[2], lvy = p0.v[3];
    float r0 = p1.v[0], r1 = p1.v[1], uox = p1.v[2], uoy = p1.v[3];
    float dpx = cu - lfx, dpy = cv - lfy, dcx = lvx, dcy = lvy, dr = r1 - r0;
    float a = (dcx * dcx + dcy * dcy) - dr * dr;
    float bq = (dpx * dcx + dpy * dcy) + r0 * dr;
    float c = (dpx * dpx + dpy * dpy) - r0 * r0;
    float discrim = bq * bq - a * c;
    vec4 color = {{0, 0, 0, 0}};
    if (discrim != 0.0f) {
        float sq = sqrtf(discrim);
        float tsx = (sq * 1.0f + bq) / a, tsy = (sq * -1.0f + bq) / a;
        if (tsx > tsy) {
            float tmp = tsx;
            tsx = tsy;
            tsy = tmp;
        }
        float t = tsx >= 0.0f ? tsx : tsy;
        color = tex(s, uox + t, uoy + 0.0f);
    }
    return color;
}

/* tile.comp:354-392 (filterBlur) */
static vec4 filter_blur(const sampler_t *s, float cu, float cv, vec4 p0, vec4 p1) {
    float sox = p0.v[0] / (float)s->w, soy = p0.v[1] / (float)s->h;
    int support = (int)p0.v[2];
    float gx = p1.v[0], gy = p1.v[1], gz = p1.v[2];
    float gauss_sum = gx;
    vec4 color = tex(s, cu, cv);
    for (int c = 0; c < 4; c++) color.v[c] *= gx;
    gx *= gy;
    gy *= gz;
    for (int i = 1; i <= support; i += 2) {
        float partial = gx;
        gx *= gy;
        gy *= gz;
        partial += gx;
        float k = (float)i + gx / partial;
        float ox = sox * k, oy = soy * k;
        vec4 a = tex(s, cu - ox, cv - oy), bq = tex(s, cu + ox, cv + oy);
        for (int c = 0; c < 4; c++) color.v[c] += (a.v[c] + bq.v[c]) * partial;
        gauss_sum += 2.0f * partial;
        gx *= gy;
        gy *= gz;
    }
    for (int c = 0; c < 4; c++) color.v[c] /= gauss_sum;
    return color;
}

/* tile.comp:394-404 (filterColorMatrix): mat4(p0 .. p3) * texel + p4, the parameters being the matrix's columns */
static vec4 filter_color_matrix(const sampler_t *s, float cu, float cv, const vec4 p[5]) {
    vec4 src = tex(s, cu, cv), r;
    for (int c = 0; c < 4; c++)
        r.v[c] = p[0].v[c] * src.v[0] + p[1].v[c] * src.v[1] + p[2].v[c] * src.v[2] + p[3].v[c] * src.v[3] + p[4].v[c];
    return r;
}

/* tile.comp:136-227 (filterText) with gamma correction off: the reference binds a 1 x 1 dummy as gamma LUT
 * (core/d3d11/renderer.cpp:262-266), so a paint that turns it on has no defined result (the product refuses it) */
static vec4 filter_text(const sampler_t *s, float cu, float cv, const vec4 p[5]) {
    const float *k = p[0].v, *bg = p[1].v, *fg = p[2].v;
    float alpha[3];
    if (k[3] == 0.0f) {
        alpha[0] = alpha[1] = alpha[2] = tex(s, cu, cv).v[0];
    } else {
        float one = 1.0f / (float)s->w, t[9];
        int wide = k[0] > 0.0f;
        for (int i = 0; i < 9; i++) /* taps -4 .. 4 (filterTextSample9Tap) */
            t[i] = ((i == 0 || i == 8) && !wide) ? 0.0f : tex(s, cu + (float)(i - 4) * one, cv).v[0];
        for (int c = 0; c < 3; c++) { /* filterTextConvolve7Tap centred on taps 3, 4, 5: dot(a0, k) + dot(a1, k.zyx) */
            const float *a = t + c;
            alpha[c] = (a[0] * k[0] + a[1] * k[1] + a[2] * k[2] + a[3] * k[3]) + (a[4] * k[2] + a[5] * k[1] + a[6] * k[0]);
        }
    }
    vec4 r;
    for (int c = 0; c < 3; c++) r.v[c] = bg[c] * (1.0f - alpha[c]) + fg[c] * alpha[c]; /* mix(bg, fg, alpha) */
    r.v[3] = 1.0f;
    return r;
}

/* tile.comp:459-562 (composite helpers) */
static float comp_div(float n, float d) { return d != 0.0f ? n / d : 0.0f; }
static void rgb_to_hsl(const float rgb[3], float hsl[3]) {
    float v = fmaxf(fmaxf(rgb[0], rgb[1]), rgb[2]), xmin = fminf(fminf(rgb[0], rgb[1]), rgb[2]);
    float c = v - xmin, l = mixf(xmin, v, 0.5f);
    float t0, t1, t2;
    if (rgb[0] == v) { t0 = 0.0f; t1 = rgb[1]; t2 = rgb[2]; }
    else if (rgb[1] == v) { t0 = 2.0f; t1 = rgb[2]; t2 = rgb[0]; }
    else { t0 = 4.0f; t1 = rgb[0]; t2 = rgb[1]; }
    hsl[0] = 1.0471975511965976f * comp_div(t0 * c + t1 - t2, c);
    hsl[1] = comp_div(c, v);
    hsl[2] = l;
}
static void hsl_to_rgb(const float hsl[3], float rgb[3]) {
    float a = hsl[1] * fminf(hsl[2], 1.0f - hsl[2]);
    const float off[3] = {0.0f, 8.0f, 4.0f};
    for (int i = 0; i < 3; i++) {
        float ks = glsl_mod(off[i] + hsl[0] * 1.9098593171027443f, 12.0f);
        rgb[i] = hsl[2] - clampf(fminf(ks - 3.0f, 9.0f - ks), -1.0f, 1.0f) * a;
    }
}
static float screen1(float d, float s) { return d + s - d * s; }
static float hard_light1(float d, float s) { return s <= 0.5f ? d * 2.0f * s : screen1(d, 2.0f * s - 1.0f); }
static float color_dodge1(float d, float s) { return d == 0.0f ? 0.0f : (s == 1.0f ? 1.0f : d / (1.0f - s)); }
static float soft_light1(float d, float s) {
    float dark = d <= 0.25f ? ((16.0f * d - 12.0f) * d + 4.0f) * d : sqrtf(d);
    float factor = s <= 0.5f ? d * (1.0f - d) : dark - d;
    return d + (s * 2.0f - 1.0f) * factor;
}
static void composite_rgb(const float d[3], const float s[3], int op, float out[3]) {
    if (op >= 0xc) {
        float dh[3], sh[3], r[3];
        rgb_to_hsl(d, dh);
        rgb_to_hsl(s, sh);
        switch (op) {
            case 0xc: r[0] = sh[0]; r[1] = dh[1]; r[2] = dh[2]; break;
            case 0xd: r[0] = dh[0]; r[1] = sh[1]; r[2] = dh[2]; break;
            case 0xe: r[0] = sh[0]; r[1] = sh[1]; r[2] = dh[2]; break;
            default: r[0] = dh[0]; r[1] = dh[1]; r[2] = sh[2]; break;
        }
        hsl_to_rgb(r, out);
        return;
    }
    for (int i = 0; i < 3; i++) {
        float dd = d[i], ss = s[i], r;
        switch (op) {
            case 0x1: r = dd * ss; break;
            case 0x2: r = screen1(dd, ss); break;
            case 0x3: r = hard_light1(ss, dd); break;
            case 0x4: r = fminf(dd, ss); break;
            case 0x5: r = fmaxf(dd, ss); break;
            case 0x6: r = color_dodge1(dd, ss); break;
            case 0x7: r = 1.0f - color_dodge1(1.0f - dd, 1.0f - ss); break;
            case 0x8: r = hard_light1(dd, ss); break;
            case 0x9: r = soft_light1(dd, ss); break;
            case 0xa: r = fabsf(dd - ss); break;
            case 0xb: r = dd + ss - 2.0f * dd * ss; break;
            default: r = ss; break;
        }
        out[i] = r;
    }
}

/* tile.comp:737-850 (main) for every framebuffer tile of one draw batch */
int pfo_frame_draw_batch(pfo_frame *f, int slot, int target_page, int color_page, uint32_t sampling_flags, int clear,
                         const float clear_color[4]) {
    batch_t *b = &f->batches[slot];
    uint8_t *target = f->dest;
    int tw = f->fb_w, th = f->fb_h;
    if (target_page >= 0) {
        if (target_page >= MAX_PAGES || !f->pages[target_page].px) return -1;
        target = f->pages[target_page].px;
        tw = f->pages[target_page].w;
        th = f->pages[target_page].h;
        clear = 1; /* renderer.cpp:382-386 */
    }
    static const uint8_t dummy[4] = {0, 0, 0, 0};
    sampler_t cs = {dummy, 1, 1, 0, 0, 0}; /* 1 x 1 dummy texture, renderer.cpp:390 */
    if (color_page >= 0 && color_page < MAX_PAGES && f->pages[color_page].px) {
        cs.px = f->pages[color_page].px;
        cs.w = f->pages[color_page].w;
        cs.h = f->pages[color_page].h;
        cs.repeat_u = (sampling_flags & 1) != 0;
        cs.repeat_v = (sampling_flags & 2) != 0;
        cs.nearest = (sampling_flags & 0xc) != 0;
    }
    float cc[4] = {0, 0, 0, 0};
    if (clear_color) memcpy(cc, clear_color, 16);

    for (int tyq = 0; tyq < f->fb_th; tyq++) {
        for (int txq = 0; txq < f->fb_tw; txq++) {
            size_t map = (size_t)tyq * f->fb_tw + txq;
            uint32_t lo = b->list_offsets[map], hi = b->list_offsets[map + 1];
            if (lo == hi && !clear) continue; /* tile.comp:743-744 */
            for (int py = 0; py < TILE; py++) {
                for (int pxq = 0; pxq < TILE; pxq++) {
                    int gx = txq * TILE + pxq, gy = tyq * TILE + py;
                    if (gx >= tw || gy >= th) continue; /* out-of-range imageStore is discarded */
                    uint8_t *dp = target + ((size_t)gy * tw + gx) * 4;
                    float dest[4];
                    if (clear) memcpy(dest, cc, 16);
                    else for (int c = 0; c < 4; c++) dest[c] = (float)dp[c] * (1.0f / 255.0f);
                    float fragx = (float)gx + 0.5f, fragy = (float)gy + 0.5f;
                    if (target_page < 0) { /* gl_FragCoord of the full canvas (strip origin) */
                        fragx = (float)(gx + f->org_tx * TILE) + 0.5f;
                        fragy = (float)(gy + f->org_ty * TILE) + 0.5f;
                    }
                    for (uint32_t k = lo; k < hi; k++) {
                        uint32_t ti = b->list_tiles[k];
                        const pfo_tile *t = &b->tiles[ti];
                        /* path lookup for paint + ctrl (the tile's control word, bound.comp:72) */
                        uint32_t plo = 0, phi = b->desc.path_count;
                        while (plo + 1 < phi) {
                            uint32_t mid = plo + (phi - plo) / 2;
                            if (ti < b->meta[mid].tile_offset) phi = mid; else plo = mid;
                        }
                        int color_entry = b->tpi[plo].color;
                        int tile_ctrl = b->tpi[plo].ctrl;
                        int backdrop;
                        float mask_alpha = 1.0f;
                        if (t->alpha_tile_id >= 0 && f->pixel_model == 1) { /* d3d9/tile.vert:93-153: the tile's backdrop rides along */
                            backdrop = t->backdrop_d3d9;
                        } else if (t->alpha_tile_id >= 0) { /* tile.comp:775-777 */
                            backdrop = 0;
                        } else {
                            backdrop = t->backdrop;
                            if (backdrop != 0 && (tile_ctrl & 0x2) && (abs(backdrop) % 2) == 0) continue; /* :786-792 */
                            tile_ctrl &= ~0x3;
                        }
                        int mask_ctrl = tile_ctrl & 0x3;
                        if (mask_ctrl != 0) { /* sampleMask, tile.comp:586-607 */
                            float cov = (f->pixel_model == 1 ? f->masks16[(size_t)t->alpha_tile_id * 256 + py * 16 + pxq]
                                                             : (float)f->masks[(size_t)t->alpha_tile_id * 256 + py * 16 + pxq] *
                                                                   (1.0f / 255.0f)) + (float)backdrop;
                            if (mask_ctrl & 0x1) cov = fabsf(cov);
                            else cov = 1.0f - fabsf(1.0f - glsl_mod(cov, 2.0f));
                            mask_alpha = mask_alpha < cov ? mask_alpha : cov;
                        }
                        /* computeTileVaryings, tile.comp:694-726 */
                        vec4 m0 = metadata_texel(f, color_entry, 0), m1 = metadata_texel(f, color_entry, 1);
                        vec4 base = metadata_texel(f, color_entry, 2);
                        vec4 fp0 = metadata_texel(f, color_entry, 3), fp1 = metadata_texel(f, color_entry, 4);
                        vec4 extra = metadata_texel(f, color_entry, 8);
                        /* mat2(colorTexMatrix0) * position + offsets: column-major (x,y)=(m0.xy), (z,w) second column */
                        float cu = m0.v[0] * fragx + m0.v[2] * fragy + m1.v[0];
                        float cv = m0.v[1] * fragx + m0.v[3] * fragy + m1.v[1];
                        int ctrl = (int)extra.v[0];
                        /* calculateColor, tile.comp:611-675 */
                        float color[4] = {base.v[0], base.v[1], base.v[2], base.v[3]};
                        int combine = (ctrl >> 8) & 0x3;
                        if (combine != 0) {
                            int filter = (ctrl >> 4) & 0xf;
                            vec4 c0;
                            if (filter == 0x1) c0 = filter_radial(&cs, cu, cv, fp0, fp1);
                            else if (filter == 0x3) c0 = filter_blur(&cs, cu, cv, fp0, fp1);
                            else if (filter == 0x2 || filter == 0x4) { /* never emitted host-side (palette.cpp:65-67) */
                                vec4 fp[5] = {fp0, fp1, metadata_texel(f, color_entry, 5), metadata_texel(f, color_entry, 6),
                                              metadata_texel(f, color_entry, 7)};
                                c0 = filter == 0x2 ? filter_text(&cs, cu, cv, fp) : filter_color_matrix(&cs, cu, cv, fp);
                            } else c0 = tex(&cs, cu, cv);
                            if (combine == 0x1) { /* SRC_IN: vec4(src.rgb, src.a * dest.a) with dest = base colour */
                                float a = c0.v[3] * color[3];
                                color[0] = c0.v[0]; color[1] = c0.v[1]; color[2] = c0.v[2]; color[3] = a;
                            } else if (combine == 0x2) { /* DEST_IN */
                                color[3] = c0.v[3] * color[3];
                            }
                        }
                        color[3] *= mask_alpha;
                        int op = (ctrl >> 10) & 0xf;
                        if (op != 0) { /* composite(), tile.comp:564-582; dest texture = the colour texture (FIXME there) */
                            vec4 dc = tex(&cs, fragx / (float)tw, fragy / (float)th);
                            float blended[3];
                            composite_rgb(dc.v, color, op, blended);
                            float sa = color[3], da = dc.v[3];
                            for (int c = 0; c < 3; c++)
                                color[c] = sa * (1.0f - da) * color[c] + sa * da * blended[c] + (1.0f - sa) * dc.v[c];
                            color[3] = 1.0f;
                        }
                        color[0] *= color[3];
                        color[1] *= color[3];
                        color[2] *= color[3];
                        for (int c = 0; c < 4; c++) dest[c] = dest[c] * (1.0f - color[3]) + color[c]; /* :841 */
                        if (f->pixel_model == 1) /* the blender writes RGBA8 after every primitive */
                            for (int c = 0; c < 4; c++) dest[c] = rintf(clampf(dest[c], 0.0f, 1.0f) * 255.0f) * (1.0f / 255.0f);
                    }
                    for (int c = 0; c < 4; c++) dp[c] = (uint8_t)rintf(clampf(dest[c], 0.0f, 1.0f) * 255.0f);
                }
            }
        }
    }
    return 0;
}

void pfo_frame_set_pixel_model(pfo_frame *f, int model) { f->pixel_model = model; }

void pfo_frame_set_origin(pfo_frame *f, int tile_x0, int tile_y0) {
    f->org_tx = tile_x0;
    f->org_ty = tile_y0;
}

void pfo_frame_pixels(const pfo_frame *f, uint8_t *out) { memcpy(out, f->dest, (size_t)f->fb_w * f->fb_h * 4); }

int pfo_frame_page_pixels(const pfo_frame *f, uint32_t page, uint8_t *out) {
    if (page >= MAX_PAGES || !f->pages[page].px) return -1;
    memcpy(out, f->pages[page].px, (size_t)f->pages[page].w * f->pages[page].h * 4);
    return 0;
}
