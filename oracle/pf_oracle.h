/* TEST INFRASTRUCTURE (oracle). Not part of the product path.
 *
 * Plain-C restatement of the reference's rasterization path, used only as the CHECKER for the
 * CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline "port" leg).
 *
 *   geometry (dice / bin / backdrops) : the reference HYBRID CPU tiler, which BASELINE.json names as
 *       the bit-exact truth: core/d3d9/tiler.cpp:46-315, core/d3d9/object_builder.cpp:19-114,
 *       core/data/segment.cpp:12-139, core/data/contour.cpp:110-173.
 *   tile resolution (propagate / alpha-tile allocation / z / per-tile lists) : tiler.cpp:369-439 restated
 *       in the GPU-driven data model of shaders/d3d11/propagate.comp:95-216 and sort.comp:49-83.
 *   pixels : shaders/d3d11/fill.comp:51-154 and tile.comp:126-134,319-404,459-607,611-675,694-726,737-850.
 *
 * PINNING. The geometry + tile-resolution half is pinned against the reference itself
 * (oracle/_ref/libpfref.so, tests/test_oracle_vs_reference.py and the fixtures in tests/golden/ it
 * generated). The pixel half is pinned against the reference's OWN compute shaders: fill.comp and tile.comp, read
 * where they lie under /root/reference, rewritten to C++ syntax (oracle/ref_harness/make_shader_cpp.py) and compiled by
 * g++ over a GLSL vocabulary header (oracle/ref_harness/glsl_shim.h) into oracle/_ref/libpfshader.so, then run on the CPU
 * on this oracle's geometry (oracle/pfshader.py): masks and frames agree within 1/255 on every fixture
 * (tests/test_oracle_golden.py; tests/golden/shader_frames.npz holds the shader-rendered frames). What that does NOT pin
 * is a particular GPU's arithmetic: the shaders are evaluated in IEEE fp32 with exact bilinear weights (a GPU filters
 * with 8 fraction bits), neither of the reference's GPU back ends can run headless here.
 *
 * Inputs are exactly what the reference's SceneBuilderD3D11 hands to its renderer
 * (core/d3d11/gpu_data.h:54-205), so the same buffers feed this oracle and the CUDA path.
 */
#ifndef PF_ORACLE_H
#define PF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfo_frame pfo_frame;

/* core/d3d11/gpu_data.h:54-94 (POD, std430-compatible). */
typedef struct {
    int32_t initial_backdrop, tile_x_offset;
    uint32_t path_index;
} pfo_backdrop_info; /* 12 B */
typedef struct {
    int32_t rect[4]; /* min_x min_y max_x max_y in tiles */
    uint32_t tile_offset, path_index, z_write, clip_path_index, backdrop_offset, pad[3];
} pfo_propagate_metadata; /* 48 B */
typedef struct {
    uint32_t global_path_id, first_global_segment_index, first_batch_segment_index, pad;
} pfo_dice_metadata; /* 16 B */
typedef struct {
    int16_t tile_min_x, tile_min_y, tile_max_x, tile_max_y;
    uint32_t first_tile_index;
    uint16_t color;
    uint8_t ctrl;
    int8_t backdrop;
} pfo_tile_path_info; /* 16 B */

typedef struct {
    uint32_t batch_id;
    uint32_t path_count, tile_count, segment_count, column_count;
    int32_t path_source;   /* 0 draw segments, 1 clip segments */
    int32_t clip_batch_id; /* batch whose tiles clip this one, or -1 */
    const pfo_backdrop_info *backdrops;
    const pfo_propagate_metadata *propagate_metadata;
    const pfo_dice_metadata *dice_metadata;
    const pfo_tile_path_info *tile_path_info;
    float transform[6]; /* m11 m21 m12 m22 tx ty (common/math/transform2.h) */
} pfo_batch_desc;

/* One flattened line (output of the dice stage), before view-box clipping. */
typedef struct {
    float from_x, from_y, to_x, to_y;
    uint32_t path_index;
} pfo_line; /* 20 B */

/* One fill, keyed by the dense (batch-local) tile index instead of the hybrid path's alpha tile id. */
typedef struct {
    uint32_t tile_index;
    uint16_t from_x, from_y, to_x, to_y;
} pfo_fill; /* 12 B */

/* Per dense tile, after propagation. */
typedef struct {
    int32_t alpha_tile_id;      /* frame-global mask slot, the clip's slot for solid-draw x alpha-clip, or -1 */
    int32_t clip_alpha_tile_id; /* mask slot min()-ed into this tile's mask, or -1 */
    int32_t fill_count;
    int8_t backdrop;       /* GPU-driven semantics (propagate.comp:129,186-188) */
    int8_t backdrop_delta; /* x-crossing sum inside the tile (object_builder.cpp:112-113) */
    int8_t backdrop_d3d9;  /* what tiler.cpp:433-434 stores (0 when clip-combined, clip backdrop when replaced) */
    uint8_t listed;        /* in its framebuffer tile's list (propagate.comp:209) */
} pfo_tile; /* 16 B */

pfo_frame *pfo_frame_create(int fb_width, int fb_height, const float view_box[4], const uint8_t *area_lut_rgba,
                            int lut_w, int lut_h);
void pfo_frame_destroy(pfo_frame *f);
/* The framebuffer shows the scene from tile (tile_x0, tile_y0) on: one horizontal strip of a larger canvas. The view
 * box stays in scene coordinates ([0, 16 * tile_y0, W, 16 * tile_y0 + fb_height]), so every float operation is the one
 * the full-canvas frame performs. Default (0, 0). */
/* 0: GPU-driven pixel semantics (fill.comp + tile.comp, default); 1: the hybrid raster shaders' (shaders/d3d9): see pf_oracle.c */
void pfo_frame_set_pixel_model(pfo_frame *f, int model);
void pfo_frame_set_origin(pfo_frame *f, int tile_x0, int tile_y0);

/* which: 0 draw, 1 clip. points = xy pairs; indices = (first_point_index, flag) pairs. */
void pfo_frame_set_segments(pfo_frame *f, int which, const float *points, uint32_t n_points,
                            const uint32_t *indices, uint32_t n_segments);
/* RGBA16F metadata texture rows (1280 texels = 128 paints x 10 texels per row), core/renderer.cpp:167-251. */
void pfo_frame_set_metadata(pfo_frame *f, const uint16_t *half_texels, uint32_t n_rows);
void pfo_frame_set_page(pfo_frame *f, uint32_t page, int w, int h, const uint8_t *rgba);

/* dice + bin + propagate + fill for one batch. Clip batches must be prepared before the batches they clip.
 * Returns a batch slot (>= 0) or a negative error. */
int pfo_frame_prepare_batch(pfo_frame *f, const pfo_batch_desc *desc);

/* counts: [0] lines [1] fills [2] alpha tiles allocated by this batch [3] first alpha tile id
 *         [4] listed tiles [5] listed after z-cull [6] max list length. */
void pfo_batch_counts(const pfo_frame *f, int slot, uint32_t counts[8]);
size_t pfo_batch_lines(const pfo_frame *f, int slot, pfo_line *out);
/* The same lines after the view-box clip of tiler.cpp:138-153 (lines entirely outside are dropped): the CUDA dice
 * stage emits these. */
size_t pfo_batch_clipped_lines(const pfo_frame *f, int slot, pfo_line *out);
/* Canonical order: by tile_index, then (from_x, from_y, to_x, to_y). */
size_t pfo_batch_fills(const pfo_frame *f, int slot, pfo_fill *out);
size_t pfo_batch_tiles(const pfo_frame *f, int slot, pfo_tile *out);
/* Per framebuffer tile: propagate.comp's z (max dense tile index of an occluding solid tile, 0 if none) and
 * the hybrid builder's z (max draw path id, core/d3d9/scene_builder.cpp:65-78). Either may be NULL. */
size_t pfo_batch_z(const pfo_frame *f, int slot, int32_t *z_d3d11, uint32_t *z_d3d9);
/* Column backdrops as bin left them (crossings above each path's tile rect), i.e. propagate's input. */
size_t pfo_batch_column_backdrops(const pfo_frame *f, int slot, int32_t *out);
/* Sorted, z-culled per-framebuffer-tile lists in CSR form (sort.comp:49-83). offsets has fb_tiles + 1 entries. */
size_t pfo_batch_tile_lists(const pfo_frame *f, int slot, uint32_t *offsets, uint32_t *dense_tile_indices);
/* 16 x 16 coverage bytes (row-major) of one mask slot. */
int pfo_frame_mask(const pfo_frame *f, uint32_t alpha_tile_id, uint8_t out[256]);

/* tile.comp for one draw batch into the destination (target_page < 0) or into a pattern page.
 * sampling_flags: core/paint/texture_allocator.h TextureSamplingFlags (REPEAT_U 1, REPEAT_V 2, NEAREST_MIN 4,
 * NEAREST_MAG 8). clear != 0 selects LOAD_ACTION_CLEAR with clear_color. */
int pfo_frame_draw_batch(pfo_frame *f, int slot, int target_page, int color_page, uint32_t sampling_flags,
                         int clear, const float clear_color[4]);
/* RGBA8, fb_width x fb_height, row-major. */
void pfo_frame_pixels(const pfo_frame *f, uint8_t *out);
int pfo_frame_page_pixels(const pfo_frame *f, uint32_t page, uint8_t *out);

/* Geometry only (dice + bin + propagate, no masks / pixels): the CPU-baseline "port" leg. */
int pfo_frame_prepare_batch_geometry_only(pfo_frame *f, const pfo_batch_desc *desc);
/* Forget all prepared batches and mask slots (start of a new frame). */
void pfo_frame_reset(pfo_frame *f);

#ifdef __cplusplus
}
#endif
#endif
