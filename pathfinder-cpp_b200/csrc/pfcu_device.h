// Internal device-side contract between the C-ABI (pfcu_api.cu) and the kernels
// (pfcu_geom.cu: dice / bin, compiled with -fmad=false for bit-exactness against the reference's x86 tiler;
//  pfcu_tiles.cu: init / scan / scatter / propagate / list building; pfcu_raster.cu: fill / composite).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/pfcu.h"

namespace pfcu {

constexpr int TILE = 16;
constexpr int GROUP_TILES = 16;  // consecutive framebuffer tiles rendered by one CTA of the tile kernel
constexpr int FB_COUNT_BITS = 20;  // FbTile::count while propagate runs (ordered tile groups): entries | masked entries << 20
constexpr uint32_t CURVE_IS_QUADRATIC = 0x80000000u;  // pathfinder/core/data/data.h:19
constexpr uint32_t CURVE_IS_CUBIC = 0x40000000u;      // pathfinder/core/data/data.h:20
constexpr float FLATTENING_TOLERANCE = 1.0f;          // pathfinder/core/d3d9/tiler.cpp:15
constexpr float FLOAT_EPSILON = 0.0001f;              // pathfinder/common/math/basic.h:11
// The reference recursion / tile walk are unbounded; the oracle (oracle/pf_oracle.c) uses the same bounds.
constexpr int MAX_FLATTEN_DEPTH = 24;
constexpr int MAX_DDA_STEPS = 65536;

// Device counters of one batch (one 64-byte line).
struct BatchCounters {
    uint32_t n_lines;         // clipped lines emitted by dice (may exceed capacity: overflow)
    uint32_t n_fills;         // total fills (from the scan)
    uint32_t n_staging;       // staging slots reserved by dice (upper bound of fills per line, summed)
    uint32_t first_alpha;     // frame-global id of this batch's first mask slot
    uint32_t n_alpha;         // mask slots of this batch = dense tiles with at least one fill (from the scan)
    uint32_t overflow;        // OverflowBits
    uint32_t scan_ticket[2];  // dynamic tile ids for the two look-back scans
    uint32_t n_list_entries;  // total list entries (from the scan over framebuffer tiles)
    uint32_t n_long;          // lines queued for the long-line bin kernel
    uint32_t n_listed;        // list entries that survive the z-cull (placed by the list scatter)
    uint32_t max_list_len;    // longest list after the z-cull
    uint32_t pad[4];
};
static_assert(sizeof(BatchCounters) == 64, "BatchCounters");

enum OverflowBits { OVF_LINES = 1, OVF_FILLS = 2, OVF_ALPHA = 4, OVF_LIST = 8, OVF_STAGING = 16, OVF_DDA = 32 };

// Per-dense-tile state written by propagate (one 16-byte store).
//   alpha     : mask slot (frame-global; the clip's slot for solid-draw x alpha-clip) or -1
//   packed    : bits 0-7 backdrop, 8-15 backdrop delta, 16-23 hybrid-equivalent backdrop, 24 listed, 25 owns a mask,
//               26-27 mask ctrl bits of the path (winding / even-odd)
//   path      : batch-local path index (what bound.comp:41-53 finds by binary search)
//   clip_alpha: mask slot min()-ed into this tile's coverage, or -1
struct TileState {
    int32_t alpha;
    uint32_t packed;
    uint32_t path;
    int32_t clip_alpha;
};

// One fill waiting for its final position (bin -> scan -> scatter). tile == ~0u marks an unused slot.
struct StagedFill {
    uint32_t tile, from, to, pad;
};

// One entry of a framebuffer tile's list: everything tile.comp reads per layer (tile.comp:765-768), one 16-byte load.
// Two formats, chosen per frame by BatchView::solid_prims:
//   general  : key, mask slot, paint | ctrl << 16 | backdrop << 24, 0  -- the tile kernel looks the paint up
//   resolved : key | LayerFlags << 24, mask slot, base colour as four halfs (r | g << 16, b | a << 16) -- every paint of
//              the frame is a plain colour (ctrl == 0): the list scatter resolves the layer completely, the tile kernel
//              reads nothing but this record (the metadata texels ARE halfs, core/renderer.cpp:176, so this is lossless)
struct TilePrim {
    uint32_t key;        // dense tile index (< 2^24, pfcu_prepare_batch checks): sort key == paint order
    int32_t alpha;       // mask slot or -1
    uint32_t ctrl_word;  // general: paint | ctrl << 16 | backdrop << 24; resolved: r | g << 16
    uint32_t pad;        // resolved: b | a << 16
};

enum LayerFlags : uint32_t {
    LF_TEXTURED = 1,  // the paint is not a plain colour (gradient, image, blur, blend mode): per-pixel shading
    LF_MASKED = 2,    // coverage comes from a mask
    LF_SKIP = 4,      // solid tile of an even-odd path with an even backdrop: invisible (tile.comp:786-792)
    LF_HEAVY = 8,     // blur filter: thousands of instructions per pixel -- the tile is split over the CTA's warps
    LF_EVEN_ODD = 16  // the mask is folded 1 - |1 - mod(c, 2)| when it is sampled (tile.comp:601-602)
};

// tile.comp:765-792 for one list entry: which of the cases above apply (paint-independent part)
__host__ __device__ __forceinline__ uint32_t layer_flags(int alpha, uint32_t ctrl, int backdrop, uint32_t mask_capacity) {
    if (alpha >= 0) {
        if (!(ctrl & 0x3u) || (uint32_t)alpha >= mask_capacity) return 0u;
        return (uint32_t)LF_MASKED | ((ctrl & 0x1u) ? 0u : (uint32_t)LF_EVEN_ODD);
    }
    const int ab = backdrop < 0 ? -backdrop : backdrop;
    return (backdrop != 0 && (ctrl & 0x2u) && (ab & 1) == 0) ? (uint32_t)LF_SKIP : 0u;
}

// Per framebuffer tile: list range + z (one 16-byte load in the composite kernel).
struct FbTile {
    uint32_t begin;   // from the scan
    uint32_t count;   // slots: entries before z-cull (counted by propagate)
    int32_t z;        // largest dense tile index of an occluding solid tile (propagate.comp:204-206)
    uint32_t cursor;  // entries actually placed by the list scatter (it leaves out what z culls, sort.comp:62)
};

struct AlphaTile {
    uint32_t tile_index;  // dense tile that owns the mask (0x7fffffff: skip) | winding rule << 31
    int32_t clip_alpha;   // mask slot to min() with, or -1
    uint32_t fill_begin;  // CSR range of its fills
    uint32_t packed;      // bits 0-7 backdrop, 8-31 fill count
};

// Per-paint constants decoded once per metadata upload from the RGBA16F texels (tile.comp:694-726).
struct Paint {
    float4 base;  // base colour
    float4 m0;    // colour texture matrix (m11 m21 m12 m22)
    float4 m1;    // xy: colour texture offset
    float4 fp0, fp1;
    int32_t ctrl;  // composite << 10 | combine << 8 | filter << 4
    int32_t pad[3];
    float4 fp2, fp3, fp4;  // filter parameters 2 - 4: only the text and colour-matrix filters read them (from the table)
};
static_assert(sizeof(Paint) == 144, "Paint");

// Everything a kernel needs to know about one batch. Passed by value.
struct BatchView {
    // inputs (device copies of the host vectors)
    const pfcu_backdrop_info *backdrops;
    const pfcu_propagate_metadata *meta;
    const pfcu_dice_metadata *dice;
    const pfcu_tile_path_info *tpi;
    const float2 *points;
    const uint2 *indices;
    uint32_t n_points, n_segments_total;
    uint32_t path_count, tile_count, segment_count, column_count;
    float transform[6];
    int identity_transform;
    float view_box[4];
    int fb_tw, fb_th;
    int fb_tx0, fb_ty0;  // scene tile shown at framebuffer tile (0, 0) (pfcu_set_target_origin: strips of a large canvas)
    // working set
    BatchCounters *counters;
    uint32_t *tile_word;    // [tile_count] low 24 bits: fill count; high 8 bits: backdrop delta
    uint32_t *fill_begin;   // [tile_count] exclusive fill offsets (CSR; final after the scan)
    uint32_t *fill_cursor;  // [tile_count] the same offsets, advanced to the end offsets by the fill scatter
    int32_t *col_backdrop;  // [column_count]
    TileState *tile_state;  // [tile_count]
    float4 *lines;          // [line_capacity] clipped lines
    uint2 *line_meta;       // [line_capacity] x: path index, y: first staging slot (long lines only)
    uint32_t *long_lines;   // [line_capacity] indices of the lines whose tile walk is shared by a warp
    uint32_t line_capacity;
    StagedFill *staging;    // [staging_capacity]
    uint32_t staging_capacity;
    uint2 *fills;           // [fill_capacity] x = from_x | from_y << 16, y = to_x | to_y << 16
    uint32_t fill_capacity;
    uint32_t *alpha_rank;   // [tile_count] exclusive count of tiles with fills (mask slot = first_alpha + rank)
    FbTile *fb;             // [fb tiles]
    TilePrim *prims;        // [prim_capacity]
    uint32_t prim_capacity;
    AlphaTile *alpha_tiles; // [alpha_capacity] batch-local
    uint32_t *alpha_map;    // [alpha_capacity] framebuffer tile of the mask's owner (~0: none), for fill's z-cull
    int cull_fill;          // fill skips masks whose tile the z-buffer culls (draw batches; fill.comp rasterizes them all)
    int fused_fill;         // the tile kernel rasterizes this batch's masks itself; k_fill is not launched (PFCU_OPT_FUSED_FILL)
    uint32_t alpha_capacity;
    unsigned long long *scan_desc[2];  // look-back descriptors (tile_count / fb tiles)
    // clip batch (may be null)
    const pfcu_propagate_metadata *clip_meta;
    const TileState *clip_tile_state;
    uint32_t clip_path_count;
    // frame-global
    uint32_t *frame_alpha_counter;  // next free mask slot
    uint8_t *masks;                 // 256 B per slot, lane-major: byte lane * 8 + q = pixel (column (lane & 3) * 4 +
                                    // (q & 3), row (lane >> 2) * 2 + (q >> 2)), so a warp stores / loads a mask with
                                    // one 8-byte access per lane and owns the same pixels as 16-byte framebuffer pieces
    uint32_t mask_capacity;         // slots
    // paints (the list scatter resolves plain colours into the list entries)
    const struct Paint *paints;
    uint32_t n_paints;
    int solid_prims;                // every paint of the frame is a plain colour: list entries carry the colour (TilePrim)
    // incremental frames (PFCU_OPT_INCREMENTAL_DICE): the first n_static_lines lines are the previous frames' dice output;
    // those whose path is marked in dirty_paths are skipped by bin, and only the segments of dice_ranges are diced again
    const uint2 *dice_ranges;       // [n_dice_ranges + 1] x: first batch segment of the range, y: segments before it; the
                                    // last entry is a sentinel (y = n_dice_segments). null: every segment is diced
    uint32_t n_dice_ranges, n_dice_segments;
    const uint32_t *dirty_paths;    // bitmap over the batch's paths, or null
    uint32_t n_static_lines;
    // Tile groups in order of cost (PFCU_OPT_ORDER_TILE_GROUPS; fb_sorted == null when off). The tile kernel's CTAs differ 14 x in
    // duration (a group of one-colour tiles against a group of 16 tiles with deep masked lists) and are placed in grid
    // order: expensive groups late in the grid leave most of the GPU idle for the last quarter of the kernel. propagate
    // counts the masked entries of every framebuffer tile (high bits of FbTile::count, FB_COUNT_BITS); the scan over
    // framebuffer tiles adds them up per group and partitions the groups -- those with at
    // least group_cost_min masked tiles first, in grid order, the others from the back -- and writes the list headers in
    // THAT order, so CTA j finds the headers of its group at j * GROUP_TILES (no dependent load).
    uint32_t group_cost_min;
    uint32_t *group_of;    // [groups] group rendered by CTA j
    uint32_t *slot_of;     // [groups] inverse: CTA that renders group g
    FbTile *fb_sorted;     // [groups * GROUP_TILES] headers in CTA order: begin / count / z from the scan, cursor from the list scatter
    // device-side timeline (pfcu_set_timeline): word 0 = records claimed so far, records from byte 64 on; null = off
    uint32_t *timeline;
    uint32_t timeline_capacity;
#ifdef PFCU_PAD_VIEW
    uint32_t pad_experiment[PFCU_PAD_VIEW / 4];
#endif
};

// ---- per-kernel argument blocks. BatchView (472 bytes) is what the host keeps; a kernel only gets the fields it reads.
// Kernel parameters are not free: padding BatchView by 64 / 256 bytes made one tiger 4096^2 frame (11 kernels) 1.0 / 2.5 us
// slower alone and 0.7 / 2.7 us slower per frame with four frames in flight (profiles/r02_tile_kernel.md section 9), i.e.
// about 1 us per 700 bytes of parameters launched. Same field names as BatchView, so the kernels read the same; the
// helpers the kernels share are templates over the block type.
#define PFCU_ARG_DECL(f) decltype(BatchView::f) f;
#define PFCU_ARG_COPY(f) memcpy(&f, &v.f, sizeof(f));
#define PFCU_ARGS(Name, FIELDS)                                           \
    struct Name {                                                         \
        FIELDS(PFCU_ARG_DECL)                                             \
        Name() = default;                                                 \
        explicit Name(const BatchView &v) { FIELDS(PFCU_ARG_COPY) }       \
    };
#ifdef PFCU_TIMELINE
#define PFCU_TL(X) X(timeline) X(timeline_capacity)
#else
#define PFCU_TL(X)
#endif
#define PFCU_INIT_FIELDS(X) \
    X(tile_word) X(col_backdrop) X(backdrops) X(fb) X(scan_desc) X(tile_count) X(column_count) X(fb_tw) X(fb_th) PFCU_TL(X)
#define PFCU_DICE_FIELDS(X)                                                                                              \
    X(dice) X(indices) X(points) X(dice_ranges) X(counters) X(line_meta) X(lines) X(long_lines) X(view_box) X(transform)  \
    X(identity_transform) X(n_dice_segments) X(n_dice_ranges) X(n_points) X(n_segments_total) X(path_count)             \
    X(segment_count) X(line_capacity) X(staging_capacity) PFCU_TL(X)
#define PFCU_BIN_FIELDS(X)                                                                                  \
    X(counters) X(line_meta) X(lines) X(long_lines) X(staging) X(tile_word) X(col_backdrop) X(meta) X(dirty_paths) \
    X(line_capacity) X(n_static_lines) X(segment_count) PFCU_TL(X)
#define PFCU_SCAN_FIELDS(X)                                                                                            \
    X(tile_word) X(fill_begin) X(fill_cursor) X(alpha_rank) X(scan_desc) X(counters) X(frame_alpha_counter) X(fb)        \
    X(fb_sorted) X(group_of) X(slot_of) X(tile_count) X(fb_tw) X(fb_th) X(fill_capacity) X(mask_capacity) \
    X(alpha_capacity) X(prim_capacity) X(group_cost_min) PFCU_TL(X)
#define PFCU_FILL_SCATTER_FIELDS(X) \
    X(counters) X(staging) X(fill_cursor) X(fills) X(staging_capacity) X(tile_count) X(fill_capacity) PFCU_TL(X)
#define PFCU_PROPAGATE_FIELDS(X)                                                                                       \
    X(backdrops) X(clip_meta) X(meta) X(tpi) X(alpha_map) X(alpha_rank) X(alpha_tiles) X(clip_tile_state) X(fb)         \
    X(fill_begin) X(fb_sorted) X(tile_state) X(col_backdrop) X(counters) X(tile_word) X(clip_path_count)              \
    X(alpha_capacity) X(fb_th) X(fb_tw) X(fb_tx0) X(fb_ty0) X(mask_capacity) X(column_count) PFCU_TL(X)
#define PFCU_LIST_FIELDS(X)                                                                                          \
    X(counters) X(meta) X(paints) X(prims) X(tile_state) X(tpi) X(fb) X(fb_sorted) X(slot_of) X(fb_tw)  \
    X(fb_tx0) X(fb_ty0) X(mask_capacity) X(n_paints) X(prim_capacity) X(solid_prims) X(tile_count) PFCU_TL(X)
#define PFCU_FILL_FIELDS(X)                                                                                  \
    X(alpha_map) X(alpha_tiles) X(counters) X(fb) X(fills) X(masks) X(alpha_capacity) X(cull_fill) X(mask_capacity) \
    X(fill_capacity) X(tile_count) PFCU_TL(X)
#define PFCU_COMPOSITE_FIELDS(X)                                                                                      \
    X(alpha_tiles) X(counters) X(fb_sorted) X(fb) X(slot_of) X(group_of) X(fills) X(masks) X(prims)       \
    X(alpha_capacity) X(fb_tw) X(fb_tx0) X(fb_ty0) X(fill_capacity) X(mask_capacity) X(prim_capacity) X(solid_prims)   \
    X(tile_count) PFCU_TL(X)
PFCU_ARGS(InitArgs, PFCU_INIT_FIELDS)
PFCU_ARGS(DiceArgs, PFCU_DICE_FIELDS)
PFCU_ARGS(BinArgs, PFCU_BIN_FIELDS)
PFCU_ARGS(ScanArgs, PFCU_SCAN_FIELDS)
PFCU_ARGS(FillScatterArgs, PFCU_FILL_SCATTER_FIELDS)
PFCU_ARGS(PropagateArgs, PFCU_PROPAGATE_FIELDS)
PFCU_ARGS(ListArgs, PFCU_LIST_FIELDS)
PFCU_ARGS(FillArgs, PFCU_FILL_FIELDS)
PFCU_ARGS(CompositeArgs, PFCU_COMPOSITE_FIELDS)

// Header of framebuffer tile `map`: where the scan put it (in CTA order when the groups are ordered)
template <class B>
__device__ __forceinline__ FbTile *fb_header(const B &b, uint32_t map) {
    if (!b.fb_sorted) return &b.fb[map];
    return &b.fb_sorted[__ldg(&b.slot_of[map / GROUP_TILES]) * GROUP_TILES + map % GROUP_TILES];
}

struct TargetView {
    uint8_t *pixels;  // RGBA8
    size_t pitch;     // bytes
    int width, height;
};

struct PaintView {
    const Paint *paints;
    uint32_t n_paints;
    int all_solid;             // every paint has ctrl == 0: base colour only, src-over
    const uint8_t *color_px;   // colour texture page (RGBA8) or a 1 x 1 dummy
    int color_w, color_h;
    uint32_t sampling_flags;
    const uint8_t *area_lut;   // 256 x 256 RGBA8
    int lut_w, lut_h;
    cudaTextureObject_t lut_tex;  // the same LUT behind the texture unit: texel fetch, clamp-to-edge, unorm8 -> float
    int lut_band;              // the LUT saturates outside |x - 152| <= 32 + y / 2 (checked at upload)
    int unit_range;            // every base colour component is in [0, 1]
};

// ---- launchers (each enqueues exactly one kernel on `s` and returns the CUDA status) --------------------------
cudaError_t launch_init(const BatchView &b, cudaStream_t s);
cudaError_t launch_dice(const BatchView &b, cudaStream_t s);
cudaError_t launch_bin(const BatchView &b, cudaStream_t s);       // walks of up to 12 tiles: a thread per line
cudaError_t launch_bin_long(const BatchView &b, cudaStream_t s);  // longer walks: a warp per line (independent of launch_bin)
cudaError_t launch_scan_tiles(const BatchView &b, cudaStream_t s);
cudaError_t launch_fill_scatter(const BatchView &b, cudaStream_t s);
cudaError_t launch_propagate(const BatchView &b, cudaStream_t s);
cudaError_t launch_scan_fb(const BatchView &b, cudaStream_t s);
cudaError_t launch_list_scatter(const BatchView &b, cudaStream_t s);
cudaError_t launch_fill(const BatchView &b, const PaintView &p, cudaStream_t s);
// origin != 0: the target is the destination framebuffer, whose tile (0, 0) is scene tile (fb_tx0, fb_ty0)
// (n_launched: kernels enqueued -- a batch that mixes plain and textured paints on a large target takes two passes)
cudaError_t launch_composite(const BatchView &b, const PaintView &p, const TargetView &t, int clear,
                             const float clear_color[4], int origin, int heavy_paints, cudaStream_t s, int *n_launched = nullptr);

bool timeline_compiled();  // the kernels were built with -DPFCU_TIMELINE
constexpr int MAX_DEVICES = 64;
int current_device();  // clamped to [0, MAX_DEVICES)
int sm_count();        // of the current device

// Device arrays of one pfcu_stroke_to_fill call (pfcu_stroke.cu)
struct StrokeArgs {
    const float2 *pts;
    const uint8_t *flags;
    const uint32_t *contour_first;
    const uint8_t *closed;
    const uint32_t *style_index;
    const pfcu_stroke_style *styles;
    const uint32_t *seg_first;  // [n_contours + 1] segment slots of every contour (both sides)
    uint32_t n_contours, n_slots;
    void *slots;                // [n_slots] SegSlot
    uint32_t *leaf_count, *leaf_offset;  // [n_slots]
    void *leaves;               // [total leaves] Leaf
    uint32_t *counts, *offsets; // [2 * n_contours] points per output contour
    uint32_t *total;            // the sum of the last scan
    uint32_t *scratch;          // scan scratch: max(n_slots, 2 * n_contours) / 4096 + 1 words
};
cudaError_t launch_stroke_segments(const StrokeArgs &a, cudaStream_t s);  // segments, leaf counts, leaf offsets (total = leaves)
cudaError_t launch_stroke_count(const StrokeArgs &a, cudaStream_t s);     // leaves, point counts, point offsets (total = points)
cudaError_t launch_stroke_write(const StrokeArgs &a, float2 *out_pts, uint8_t *out_flags, cudaStream_t s);
size_t stroke_slot_bytes();
size_t stroke_leaf_bytes();

// Programmatic dependent launch: a kernel launched with launch_pdl may be scheduled while the previous kernel of its
// stream is still draining (its CTAs are placed and run up to pdl_wait(), which returns once that kernel has completed
// and its writes are visible), so the launch latency between the small kernels of a frame overlaps the previous kernel's
// tail. Every kernel calls pdl_wait() before it touches anything another kernel produced. PFCU_NO_PDL builds use plain
// launches (the wait is then a no-op).
__device__ __forceinline__ void pdl_wait() {
#if defined(__CUDA_ARCH__)
    asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
// griddepcontrol.launch_dependents right after this wait (so that the next kernel's CTAs are placed while this grid runs and
// sit in their own pdl_wait()) was measured and is NOT used: the gap between the last CTA of a kernel and the first
// running CTA of the next stays 1.3 - 3.4 us (tools/timeline.py; most edges of the frame graph also carry a cross-stream
// dependency, which is a full edge), one frame alone takes 100.9 instead of 102.4 us, and the waiting CTAs hold registers
// and shared memory that the other frames in flight would use: 65.4 instead of 60.9 us per frame with 4 contexts.
// Device-side timeline: thread 0 of every CTA records when the CTA was placed, when the kernel it depends on had
// finished (griddepcontrol.wait returned) and when the thread left the kernel, with the SM it ran on -- %globaltimer is
// one clock for the whole device, so the records of all kernels, streams and contexts line up (what neither CUDA events
// nor ncu, which serialises kernels, can show: which kernels of which frames actually share the GPU).
struct CtaStamp {
    uint32_t *tl;
    uint32_t cap, stage;
    unsigned long long *t;  // shared memory: placed, started (keeps the stamps out of the kernel's registers)
    static __device__ __forceinline__ unsigned long long now() {
        unsigned long long v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
        return v;
    }
    __device__ __forceinline__ void placed() const {
        if (tl && threadIdx.x == 0) t[0] = now();
    }
    __device__ __forceinline__ void go() const {
        if (tl && threadIdx.x == 0) t[1] = now();
    }
    __device__ __forceinline__ ~CtaStamp() {
        if (tl && threadIdx.x == 0) {
            const unsigned long long t_end = now();
            const uint32_t i = atomicAdd(tl, 1u);
            if (i < cap) {
                uint32_t sm;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
                pfcu_timeline_record *r = reinterpret_cast<pfcu_timeline_record *>(reinterpret_cast<char *>(tl) + 64) + i;
                r->stage = stage;
                r->sm = sm;
                r->cta = blockIdx.x;
                r->n_ctas = gridDim.x;
                r->t_placed_ns = t[0];
                r->t_start_ns = t[1];
                r->t_end_ns = t_end;
            }
        }
    }
};
// first statement of every kernel of the frame: stamp, wait for the kernel this one depends on, stamp
// The stamps are compiled in only with -DPFCU_TIMELINE (lib/libpfcu_trace.so): with the timeline off they still cost a
// frame of eleven kernels 1.8 us alone and 1.3 us per frame in a stream of frames.
#ifndef PFCU_TIMELINE
#define PFCU_KERNEL_BEGIN(b, stage_) pdl_wait()
#else
#define PFCU_KERNEL_BEGIN(b, stage_)                                                                      \
    __shared__ unsigned long long cta_stamp_t_[2];                                                        \
    const CtaStamp cta_stamp_{(b).timeline, (b).timeline_capacity, (uint32_t)(stage_), cta_stamp_t_};     \
    cta_stamp_.placed();                                                                                  \
    pdl_wait();                                                                                           \
    cta_stamp_.go()
#endif

// (One shared-memory carve-out for every kernel -- an SM cannot change its L1 / shared-memory split while CTAs are resident,
// and the timeline shows fill getting no CTA on the 32 SMs that hold a CTA of the scan over framebuffer tiles until that
// has finished -- was measured: cudaFuncAttributePreferredSharedMemoryCarveout = 72 % on all kernels, or only on fill and
// the small kernels beside it, gains 0.6 - 0.8 us per frame in a stream of frames and loses 1.2 - 1.4 us on a frame
// alone, the kernels that live on L1 hits paying for it; left to the driver. profiles/r02_tile_kernel.md section 10.)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, 1, 1);
    cfg.blockDim = dim3(block, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
#ifndef PFCU_NO_PDL
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#else
    (void)attr;
#endif
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

}  // namespace pfcu
