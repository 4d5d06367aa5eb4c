/* pfcu -- C ABI of the B200-native GPU-driven rasterization path (bound, dice, bin, propagate, sort, fill, tile).
 *
 * This is the drop-in boundary. It replaces, for this path only, what floppyhammer/pathfinder-cpp's
 * RendererD3D11 (pathfinder/core/d3d11/renderer.cpp:113-1079) reaches through the abstract GPU layer
 * (pathfinder/gpu/device.h:24-147, command_encoder.h:158-273, queue.h:11-25), the GpuMemoryAllocator
 * (pathfinder/gpu_mem/allocator.h:55-108) and the seven compute shaders (pathfinder/shaders/d3d11/[name].comp).
 * Inputs are exactly the vectors SceneBuilderD3D11 produces (pathfinder/core/d3d11/scene_builder.h:50-55,
 * gpu_data.h:54-205); the host adapter that calls these entry points from the reference's own Renderer
 * interface is pathfinder-cpp_b200/host/renderer_cuda.{h,cpp}; INTEGRATION.md shows the wiring.
 *
 * Conventions: plain pointers and sizes, no C++/torch types. Every call returns PFCU_OK (0) or a negative
 * pfcu_status and records a message retrievable with pfcu_last_error() (thread-local). All work is enqueued
 * on ONE CUDA stream per context (pfcu_set_stream; default: a private non-blocking stream). Host buffers
 * passed to upload calls are consumed before the call returns (staged through pinned memory), so the caller
 * may free them immediately -- the same lifetime rule as CommandEncoder::write_buffer
 * (pathfinder/gpu/command_encoder.cpp:205-239). There is NO CPU fallback: without a CUDA device every entry
 * point fails with PFCU_ERR_CUDA.
 */
#ifndef PFCU_H
#define PFCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFCU_ABI_VERSION 1

typedef enum {
    PFCU_OK = 0,
    PFCU_ERR_INVALID = -1,  /* bad argument / unknown batch or page (reference: Logger::error + early return) */
    PFCU_ERR_CUDA = -2,     /* CUDA runtime error, no device */
    PFCU_ERR_OOM = -3,      /* device allocation failed */
    PFCU_ERR_OVERFLOW = -4, /* a stage ran out of space twice (reference: "Ran out of space ...", renderer.cpp:551,575) */
    PFCU_ERR_STATE = -5     /* call sequence violated (e.g. draw before prepare) */
} pfcu_status;

typedef struct pfcu_ctx pfcu_ctx;

/* ---- records shared with the reference host code (pathfinder/core/d3d11/gpu_data.h), POD, std430-compatible */
typedef struct { /* BackdropInfoD3D11, gpu_data.h:54-60 */
    int32_t initial_backdrop, tile_x_offset;
    uint32_t path_index;
} pfcu_backdrop_info;
typedef struct { /* PropagateMetadataD3D11, gpu_data.h:62-74 */
    int32_t tile_rect[4];
    uint32_t tile_offset, path_index, z_write, clip_path_index, backdrop_offset, pad0, pad1, pad2;
} pfcu_propagate_metadata;
typedef struct { /* DiceMetadataD3D11, gpu_data.h:76-82 */
    uint32_t global_path_id, first_global_segment_index, first_batch_segment_index, pad;
} pfcu_dice_metadata;
typedef struct { /* TilePathInfoD3D11, gpu_data.h:84-94 */
    int16_t tile_min_x, tile_min_y, tile_max_x, tile_max_y;
    uint32_t first_tile_index;
    uint16_t color;
    uint8_t ctrl;
    int8_t backdrop;
} pfcu_tile_path_info;

/* TileBatchDataD3D11 + PrepareTilesInfoD3D11 (gpu_data.h:97-165) flattened. */
typedef struct {
    uint32_t batch_id;
    uint32_t path_count, tile_count, segment_count, column_count;
    int32_t path_source;   /* 0 = draw segments, 1 = clip segments (PathSource) */
    int32_t clip_batch_id; /* ClippedPathInfo::clip_batch_id or -1 */
    const pfcu_backdrop_info *backdrops;               /* column_count */
    const pfcu_propagate_metadata *propagate_metadata; /* path_count */
    const pfcu_dice_metadata *dice_metadata;           /* path_count */
    const pfcu_tile_path_info *tile_path_info;         /* path_count */
    float transform[6];                                /* m11 m21 m12 m22 m13 m23 (Transform2) */
} pfcu_batch_desc;

/* ---- parity taps (device results copied to the host) */
typedef struct {
    float from_x, from_y, to_x, to_y;
    uint32_t path_index;
} pfcu_line; /* 20 B: one flattened line after the view-box clip, output of dice */
typedef struct {
    uint32_t tile_index; /* dense, batch-local */
    uint16_t from_x, from_y, to_x, to_y;
} pfcu_fill; /* 12 B */
typedef struct {
    int32_t alpha_tile_id;      /* frame-global mask slot (the clip's for solid-draw x alpha-clip) or -1 */
    int32_t clip_alpha_tile_id; /* mask slot min()-ed into this tile's mask or -1 */
    int32_t fill_count;
    int8_t backdrop;
    int8_t backdrop_delta;
    int8_t backdrop_d3d9; /* the value the reference's hybrid tiler stores for the same tile (tiler.cpp:433-434) */
    uint8_t listed;
} pfcu_tile; /* 16 B */

/* Per-frame counters (also the units of the throughput metrics). */
typedef struct {
    uint32_t batches, segments, lines, fills, alpha_tiles, dense_tiles, listed_tiles, listed_after_cull;
    uint32_t fb_tiles, max_list_len, overflow_flags, retries;
    uint32_t kernel_launches; /* kernels enqueued for this frame, including retries */
    uint32_t diced_segments; /* segments that went through dice this frame (all of them unless PFCU_OPT_INCREMENTAL_DICE) */
    uint32_t uploaded_bytes; /* host-to-device bytes of this frame: segments (full or partial uploads) + batch metadata */
    uint32_t reserved[1];
    float gpu_ms; /* CUDA-event time of the frame on the context's stream (prepare..last draw), last attempt */
} pfcu_frame_stats;

/* Stroke style, as StrokeStyle (pathfinder/core/stroke.h:14-26). line_cap: 0 butt, 1 square, 2 round (LineCap);
 * line_join: 0 miter, 1 bevel, 2 round (LineJoin). */
typedef struct {
    float line_width;
    int32_t line_cap;
    int32_t line_join;
    float miter_limit;
    /* Canvas::push_path transforms the finished outline by the canvas state's transform (core/canvas.cpp:190,
     * Outline::transform, core/data/path.cpp:7-22): m11 m21 m12 m22 m13 m23, applied to every output point; an identity
     * (1 0 0 1 0 0) leaves the points untouched, as upstream. */
    float transform[6];
} pfcu_stroke_style;

/* ---- lifecycle */
int pfcu_abi_version(void);
const char *pfcu_last_error(void);
int pfcu_create(int device_ordinal, pfcu_ctx **out);
void pfcu_destroy(pfcu_ctx *ctx);
/* Run on the caller's stream (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream). NULL restores
 * the private stream. */
int pfcu_set_stream(pfcu_ctx *ctx, void *cuda_stream);
/* The stream work is enqueued on (cudaStream_t), for event timing by the caller. */
void *pfcu_get_stream(pfcu_ctx *ctx);

/* ---- static resources (Renderer::Renderer, pathfinder/core/renderer.cpp:13-37) */
/* 256 x 256 RGBA8 area LUT as decoded by the reference from shaders/area_lut.png. */
int pfcu_set_area_lut(pfcu_ctx *ctx, const uint8_t *rgba, int width, int height);

/* ---- per-target state: Renderer::set_dest_texture + Scene view box (core/renderer.h:81, core/scene.cpp:170-179) */
/* rgba8_dev == NULL: the context owns the destination (width*4 pitch). Otherwise render into caller memory
 * (device pointer, pitch in bytes, 16-byte aligned rows). view_box = {left, top, right, bottom}. */
int pfcu_set_target(pfcu_ctx *ctx, int width, int height, void *rgba8_dev, size_t pitch_bytes,
                    const float view_box[4]);

/* The destination shows the scene from pixel (origin_x, origin_y) on (multiples of 16; reset to 0 by pfcu_set_target):
 * one horizontal strip of a larger canvas per GPU. The view box passed to pfcu_set_target stays in scene coordinates
 * ([0, origin_y, W, origin_y + height]) and the batch metadata is built against it, so every float operation is the
 * one the full-canvas frame performs: the strip is bit-identical to its rows of the full frame. (The reference has no
 * equivalent; its tile kernel maps scene tile (x, y) to framebuffer pixel (16 x, 16 y), tile.comp:739-741.) */
int pfcu_set_target_origin(pfcu_ctx *ctx, int origin_x, int origin_y);

/* ---- per-scene uploads */
/* RendererD3D11::upload_scene (d3d11/renderer.cpp:346-350): which 0 = draw, 1 = clip.
 * points: xy float pairs; indices: SegmentIndicesD3D11 {first_point_index, flag} pairs. */
int pfcu_upload_scene(pfcu_ctx *ctx, int which, const float *points, uint32_t n_points, const uint32_t *indices,
                      uint32_t n_segments);
/* Renderer::upload_texture_metadata (core/renderer.cpp:167-251): rows of 1280 RGBA16F texels. */
/* Incremental scenes (replaces what Scene::epoch / LastSceneInfo::draw_segment_ranges exist for upstream, core/scene.h:32-49,
 * core/d3d11/scene_builder.cpp:217-218 -- the reference records the ranges and then re-uploads and re-dices everything): the
 * points [first_point, first_point + n_points) of source `which` changed; they belong to the segments [first_segment,
 * first_segment + n_segments) (one path's Range of draw_segment_ranges, or several consecutive paths'). Topology -- the
 * indices -- is unchanged; anything else needs pfcu_upload_scene. Only this range crosses PCIe, and with
 * PFCU_OPT_INCREMENTAL_DICE only the paths that own those segments are diced again in the next frame. */
int pfcu_update_scene_range(pfcu_ctx *ctx, int which, uint32_t first_point, const float *points, uint32_t n_points,
                            uint32_t first_segment, uint32_t n_segments);
int pfcu_upload_paint_metadata(pfcu_ctx *ctx, const uint16_t *half_texels, uint32_t n_rows);
/* Renderer::allocate_pattern_texture_page / upload_texel_data (core/renderer.cpp:46-61,95-114). */
int pfcu_alloc_page(pfcu_ctx *ctx, uint32_t page, int width, int height);
int pfcu_upload_page_region(pfcu_ctx *ctx, uint32_t page, int x, int y, int width, int height, const uint8_t *rgba);

/* ---- stroke-to-fill (replaces OutlineStrokeToFill::offset, pathfinder/core/stroke.cpp:124-167, which Canvas::stroke_path runs
 * on the CPU for every stroked path, core/canvas.cpp:272-301) for a batch of contours at once. Input, host memory: the
 * contours' points (x, y) and point flags (0 on-curve, 1 first / quadratic control point, 2 second control point:
 * PointFlag, core/data/data.h:42-51) back to back; contour i owns points [contour_first[i], contour_first[i + 1]) and is
 * closed iff closed[i]; it is stroked with styles[style_index[i]]. pfcu_stroke_to_fill runs the two passes and keeps the result
 * in the context; its contour / point totals come back through the out parameters, pfcu_stroke_result copies it out:
 * out_contour_first has n_out_contours + 1 entries. A closed contour yields two output contours (outer and inner offset),
 * an open one a single contour with its caps, in input order -- the Outline the reference would push_draw_path. Every
 * output point equals the reference's bit for bit (the kernels are compiled without FMA contraction, like the reference). */
int pfcu_stroke_to_fill(pfcu_ctx *ctx, const float *points, const uint8_t *flags, uint32_t n_points, const uint32_t *contour_first,
                        const uint8_t *closed, const uint32_t *style_index, uint32_t n_contours, const pfcu_stroke_style *styles,
                        uint32_t n_styles, uint32_t *n_out_contours, uint32_t *n_out_points);
int pfcu_stroke_result(pfcu_ctx *ctx, float *out_points, uint8_t *out_flags, uint32_t *out_contour_first);
/* Device time (ms, CUDA events) of the two passes of the last pfcu_stroke_to_fill / pfcu_dash_outlines (valid after the
 * matching _result call). */
float pfcu_stroke_gpu_ms(pfcu_ctx *ctx);
/* Dashing (replaces OutlineDash::dash + into_outline, pathfinder/core/dash.cpp:49-65, which Canvas::stroke_path runs before the
 * stroker when a line dash is set, core/canvas.cpp:286-291) for a batch of outlines. Outline o owns the contours
 * [outline_first[o], outline_first[o + 1]) of the input (same arrays as pfcu_stroke_to_fill) and is dashed with the pattern
 * dashes[dash_first[o] .. dash_first[o + 1]) at phase dash_offset[o]; the dash state runs on from one contour of an outline
 * into the next, as upstream. The result -- open contours, ready for pfcu_stroke_to_fill -- stays in the context;
 * pfcu_dash_result copies it out: out_contour_first has n_out_contours + 1 entries, out_outline_first n_outlines + 1
 * (contour ranges per input outline). Bit-identical to the reference's points. */
int pfcu_dash_outlines(pfcu_ctx *ctx, const float *points, const uint8_t *flags, uint32_t n_points, const uint32_t *contour_first,
                       const uint8_t *closed, uint32_t n_contours, const uint32_t *outline_first, uint32_t n_outlines,
                       const float *dashes, const uint32_t *dash_first, const float *dash_offset, uint32_t *n_out_contours,
                       uint32_t *n_out_points);
int pfcu_dash_result(pfcu_ctx *ctx, float *out_points, uint8_t *out_flags, uint32_t *out_contour_first, uint32_t *out_outline_first);

/* ---- frame: RendererD3D11::draw (d3d11/renderer.cpp:302-336) */
int pfcu_begin_frame(pfcu_ctx *ctx);
/* prepare_tiles (renderer.cpp:510-616): bound + dice + bin + propagate + fill + sort for one batch.
 * Clip batches must be prepared before the batches they clip (the reference submits them first, :318-327). */
int pfcu_prepare_batch(pfcu_ctx *ctx, const pfcu_batch_desc *desc);
/* draw_tiles (renderer.cpp:365-448): composite a prepared draw batch into the destination (target_page < 0) or
 * into a pattern page (render target). color_page < 0: the 1 x 1 dummy texture. sampling_flags:
 * TextureSamplingFlags (REPEAT_U 1, REPEAT_V 2, NEAREST_MIN 4, NEAREST_MAG 8). clear != 0: LOAD_ACTION_CLEAR. */
int pfcu_draw_batch(pfcu_ctx *ctx, uint32_t batch_id, int target_page, int color_page, uint32_t sampling_flags,
                    int clear, const float clear_color[4]);
/* Waits for the frame, checks the device-side capacity flags and, when a stage overflowed, grows its buffer
 * and replays the recorded frame (the reference's retry loops, renderer.cpp:537-577, without the mid-frame
 * read-backs). stats may be NULL. */
int pfcu_end_frame(pfcu_ctx *ctx, pfcu_frame_stats *stats);
/* pfcu_end_frame in two halves, for callers that keep several frames in flight (one context per frame in flight,
 * e.g. double buffering): pfcu_submit_frame enqueues the rest of the frame and its read-back and returns without
 * waiting; pfcu_wait_frame blocks until that frame is done and does the overflow check / replay / stats of
 * pfcu_end_frame. Between the two calls the context accepts no other frame or upload call (PFCU_ERR_STATE). The
 * reference submits and blocks on its one fence in the same step (Queue::submit, gpu/queue.h:17; core/renderer.cpp:30). */
int pfcu_submit_frame(pfcu_ctx *ctx);
int pfcu_wait_frame(pfcu_ctx *ctx, pfcu_frame_stats *stats);

/* ---- options */
enum {
    /* 1 (default): when two consecutive frames enqueue identical work (same batches, counts, buffers, paints, target),
     * the frame is captured as a CUDA graph and every following identical frame is ONE graph launch issued by
     * pfcu_end_frame (pfcu_prepare_batch / pfcu_draw_batch then only record). The reference has the matching notion in
     * its scene epochs (core/scene.h:32-49). A frame that differs is enqueued kernel by kernel and the graph dropped.
     * 0: always enqueue kernel by kernel. */
    PFCU_OPT_RETAIN_FRAME_GRAPH = 0,
    /* 0 (default): the fill stage skips the masks of draw tiles that the z-buffer culls (tiles under an opaque whole-tile
     * layer of a later path, sort.comp:62): nothing ever reads them (19 % of tiger.svg's masks at 4096^2). 1: every mask
     * is rasterized, as fill.comp:109-154 does (pfcu_read_mask then returns a valid mask for culled tiles too). */
    PFCU_OPT_FILL_CULLED_TILES = 1,
    /* 1: a batch keeps the lines dice produced (keyed on batch id, path / segment counts, dice metadata, transform, view box
     * and the scene upload they came from); later frames dice only the paths touched by pfcu_update_scene_range since then
     * and bin skips the retained lines of those paths. A frame that would re-dice more than half of the batch, or whose key
     * changed, dices everything and becomes the new base. Default 0: every frame dices every segment, like the reference. */
    PFCU_OPT_INCREMENTAL_DICE = 2,
    /* 1: the tile kernel rasterizes the masks of a draw batch itself, in shared memory, right before it blends them
     * (fill.comp:109-154 inside tile.comp:737-850): no separate fill launch for the batch, no mask written to or read
     * from device memory, only masks that a tile list references are rasterized. The mask bytes are computed by the
     * same code as the separate fill kernel, so the frame is byte-identical either way (tested on every fixture). Clip
     * batches always go through the separate fill kernel (other batches read their masks). 0 (default): separate fill
     * kernel for every batch -- measured faster on B200 (tiger.svg @ 4096^2: fill 24.6 us beside the list building +
     * tile 29.1 us, against 64.2 us for the fused tile kernel: the masked tiles of a 16-tile group serialise behind one
     * CTA, profiles/r02_tile_kernel.md section 7). pfcu_read_mask needs 0. */
    PFCU_OPT_FUSED_FILL = 3,
    /* n > 0 (default 4): the tile kernel renders the EXPENSIVE groups of 16 framebuffer tiles first -- propagate counts
     * the masked tiles of every group, the scan over framebuffer tiles puts the groups with at least n of them first (in
     * grid order) and the others last, and writes the list headers in that order -- instead of wherever the scene puts
     * them in the grid (a group costs 1.2 .. 17 us, and expensive groups late in the grid leave most of the GPU idle
     * at the end of the kernel; tools/timeline.py). Same pixels, same lists. 0: groups in grid order. */
    PFCU_OPT_ORDER_TILE_GROUPS = 4,
    /* 1 (default): the batches of a frame after its first are PREPARED side by side (bound .. fill of batch i on stream pair
     * i mod 4), each after the clip batch it reads; the tile passes stay in submission order on the context's stream, each waiting for its own batch.
     * The reference prepares and draws batch after batch (d3d11/renderer.cpp:318-336); a frame of many small batches -- the
     * demo's primitives scene: 14 batches, render-target passes, a blurred shadow -- is then a chain of 150 tiny kernels.
     * Demo scene at 2048^2: 1006 -> 724 us per frame alone, 690 -> 629 us streamed; a frame of one batch is enqueued as
     * before. Mask slots are handed out in completion order, so alpha tile ids can differ from run to run; pixels do not.
     * 0: batch after batch. */
    PFCU_OPT_CONCURRENT_BATCHES = 5
};
int pfcu_set_option(pfcu_ctx *ctx, int option, int value);

/* ---- measurement */
enum {
    PFCU_STAGE_INIT = 0, /* "bound" */
    PFCU_STAGE_DICE,
    PFCU_STAGE_BIN,
    PFCU_STAGE_SCAN_TILES,
    PFCU_STAGE_FILL_SCATTER,
    PFCU_STAGE_PROPAGATE,
    PFCU_STAGE_SCAN_FB,
    PFCU_STAGE_LIST_SCATTER, /* "sort", first half: contiguous z-culled list per framebuffer tile (paint order is
                                established on chip by the tile kernel) */
    PFCU_STAGE_FILL,
    PFCU_STAGE_COMPOSITE, /* "tile" */
    PFCU_NUM_STAGES
};
/* enabled != 0: bracket every kernel of subsequent frames with CUDA events on the context's stream. */
int pfcu_set_profiling(pfcu_ctx *ctx, int enabled);
/* Per-stage device time (ms, summed over batches) of the last frame ended with profiling on. n <= PFCU_NUM_STAGES. */
int pfcu_get_stage_times(pfcu_ctx *ctx, float *ms, int n);

/* ---- tracing: a device-side timeline of every CTA of every kernel of the frames rendered while it is on. Thread 0 of a
 * CTA reads %globaltimer (one nanosecond clock for the whole device) when the CTA is placed on an SM, when the kernel it
 * depends on has finished (programmatic dependent launch: CTAs are placed early and wait) and when it leaves the kernel.
 * Unlike CUDA events (one stream) and ncu (serialises kernels) this shows which kernels of which frames and contexts share
 * the GPU at any moment. Costs one predictable branch per CTA when off. */
typedef struct pfcu_timeline_record {
    uint32_t stage;  /* PFCU_STAGE_*; | 0x100: the second kernel of the stage (long-walk bin) */
    uint32_t sm;     /* %smid */
    uint32_t cta, n_ctas;
    uint64_t t_placed_ns, t_start_ns, t_end_ns;
} pfcu_timeline_record;
/* capacity > 0: (re)allocate a device buffer of that many records and record from the next frame on; 0: off. */
int pfcu_set_timeline(pfcu_ctx *ctx, uint32_t capacity);
/* Waits for the context, copies out the records written since the last call (at most max_records; NULL: count only) and
 * starts over. Returns the number of records written (may exceed the capacity: the excess was dropped), or -1. */
int64_t pfcu_read_timeline(pfcu_ctx *ctx, pfcu_timeline_record *out, int64_t max_records);

/* ---- whole-frame CUDA graph: replay the last completed frame with device-resident inputs (no host memory is
 * touched, no allocation, no read-back inside the graph). Segment points may be re-uploaded between replays
 * (animation); the batch structure must stay the same. */
int pfcu_graph_capture(pfcu_ctx *ctx);
int pfcu_graph_launch(pfcu_ctx *ctx);                          /* asynchronous, on the context's stream */
int pfcu_graph_finish(pfcu_ctx *ctx, pfcu_frame_stats *stats); /* waits, reads the counters, reports overflow */

/* ---- results */
int pfcu_read_target(pfcu_ctx *ctx, uint8_t *host_rgba8);                    /* width*height*4, tightly packed */
/* CommandEncoder::read_texture(texture, region, data) (gpu/command_encoder.cpp:317-355): width*height*4 bytes, tightly
 * packed; a region that is empty or not inside the target is PFCU_ERR_INVALID (the reference logs and returns). */
int pfcu_read_target_region(pfcu_ctx *ctx, int x, int y, int width, int height, uint8_t *host_rgba8);
int pfcu_read_page(pfcu_ctx *ctx, uint32_t page, uint8_t *host_rgba8);
/* The same read-back without blocking (replaces CommandEncoder::read_texture + the fence wait of Queue::submit_and_wait,
 * gpu/command_encoder.cpp:317-355, gpu/vk/queue.cpp:29-35): enqueues the device-to-host copy of the whole target behind the
 * frame in flight, on the context's copy stream, and returns. host_rgba8 should be page-locked (pfcu_host_alloc);
 * host_pitch_bytes 0 = tightly packed. pfcu_wait_read blocks until the pixels are in host memory. A later frame on the same
 * context waits (on the device) for the copy before it overwrites the target; other contexts keep rendering. */
int pfcu_read_target_async(pfcu_ctx *ctx, uint8_t *host_rgba8, size_t host_pitch_bytes);
int pfcu_wait_read(pfcu_ctx *ctx);
void *pfcu_host_alloc(size_t bytes); /* page-locked host memory; NULL + pfcu_last_error() on failure */
void pfcu_host_free(void *p);
void *pfcu_target_device_ptr(pfcu_ctx *ctx, size_t *pitch_bytes);            /* zero-copy consumers */

/* ---- parity taps: pass NULL to query the element count. Valid after pfcu_end_frame. */
int64_t pfcu_read_lines(pfcu_ctx *ctx, uint32_t batch_id, pfcu_line *out);
int64_t pfcu_read_fills(pfcu_ctx *ctx, uint32_t batch_id, pfcu_fill *out); /* sorted by tile, then coordinates */
int64_t pfcu_read_tiles(pfcu_ctx *ctx, uint32_t batch_id, pfcu_tile *out);
int64_t pfcu_read_z(pfcu_ctx *ctx, uint32_t batch_id, int32_t *out);       /* per framebuffer tile */
/* Sorted, z-culled per-framebuffer-tile lists, CSR: offsets has fb_tiles + 1 entries. */
int64_t pfcu_read_tile_lists(pfcu_ctx *ctx, uint32_t batch_id, uint32_t *offsets, uint32_t *dense_tile_indices);
int pfcu_read_mask(pfcu_ctx *ctx, uint32_t alpha_tile_id, uint8_t out[256]); /* 16 x 16 coverage, row-major */

#ifdef __cplusplus
}
#endif
#endif /* PFCU_H */
