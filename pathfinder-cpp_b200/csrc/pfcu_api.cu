// C-ABI (include/pfcu.h) over the kernels: context, device arena, frame recording / replay, parity taps.
//
// Replaces, for this path, the reference's RendererD3D11 host orchestration
// (pathfinder/core/d3d11/renderer.cpp:302-616) and its GpuMemoryAllocator (pathfinder/gpu_mem/allocator.cpp):
// buffers are device-local, keyed by batch slot, grown geometrically and reused across frames; the whole frame is
// enqueued on one stream without a single mid-frame host read-back (the reference has three per batch,
// renderer.cpp:705-724,832-845,943-950). Capacity overflows are detected once, at pfcu_end_frame, and answered by
// growing the buffer and replaying the recorded frame. A context is ONE frame in flight (pfcu_submit_frame /
// pfcu_wait_frame; pfcu_end_frame = both): applications that stream frames keep a few contexts.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only; a no-op unless a profiler attaches

#include "pfcu_device.h"

using namespace pfcu;

namespace pfcu {
// pfcu_stroke.cu (the stroke launchers are declared in pfcu_device.h)
cudaError_t launch_dash_count(const float2 *pts, const uint8_t *flags, const uint32_t *contour_first, const uint8_t *closed,
                              const uint32_t *outline_first, const float *dashes, const uint32_t *dash_first, const float *dash_offset,
                              uint32_t n_outlines, uint32_t *point_counts, uint32_t *contour_counts, uint32_t *point_offsets,
                              uint32_t *contour_offsets, uint32_t *totals, uint32_t *scratch, cudaStream_t s);
cudaError_t launch_dash_write(const float2 *pts, const uint8_t *flags, const uint32_t *contour_first, const uint8_t *closed,
                              const uint32_t *outline_first, const float *dashes, const uint32_t *dash_first, const float *dash_offset,
                              uint32_t n_outlines, uint32_t *point_counts, uint32_t *contour_counts, const uint32_t *point_offsets,
                              const uint32_t *contour_offsets, float2 *out_pts, uint8_t *out_flags, uint32_t *out_contour_first,
                              cudaStream_t s);
}  // namespace pfcu

namespace {

thread_local std::string g_last_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            return fail(_e == cudaErrorMemoryAllocation ? PFCU_ERR_OOM : PFCU_ERR_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(_e), __FILE__, __LINE__);                                         \
    } while (0)

size_t round_up_pow2(size_t v) {
    size_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(round_up_pow2(bytes), 256);
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T *as() const {
        return static_cast<T *>(p);
    }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        size_t want = std::max<size_t>(round_up_pow2(bytes), 4096);
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr int MAX_PAGES = 64;
constexpr int MAX_SLOTS = 256;

struct Page {
    DevBuf px;
    int w = 0, h = 0;
};

struct BatchSlot {
    pfcu_batch_desc desc{};
    PinnedBuf host_meta;  // the four metadata vectors, packed, kept for replay
    size_t off_backdrops = 0, off_meta = 0, off_dice = 0, off_tpi = 0, meta_bytes = 0;
    DevBuf dev_meta, tile_word, fill_begin, fill_cursor, alpha_rank, col_backdrop, tile_state, lines, line_meta, staging, fills, fb, long_lines,
        prims, alpha_tiles, alpha_map, scan_desc0, scan_desc1, group_of, slot_of, fb_sorted;
    uint32_t line_cap = 0, fill_cap = 0, staging_cap = 0;
    BatchView view{};
    bool prepared = false;
    bool heavy_paints = false;  // a path of the batch paints with a blur filter (tile.comp:354-392)
    // PFCU_OPT_CONCURRENT_BATCHES: the ends of the batch's two kernel chains (on its lane's streams) and whether the
    // context's stream has waited for them yet
    cudaEvent_t lane_done[2] = {nullptr, nullptr};
    bool lane_joined[2] = {true, true};
    // Incremental frames (PFCU_OPT_INCREMENTAL_DICE): the dice output of an earlier frame that this slot still holds.
    struct DiceBase {
        bool valid = false;
        uint64_t key = 0;        // everything the retained lines depend on except the segments themselves
        uint64_t scene_gen = 0;  // full upload the lines were diced from
        uint64_t seq = 0;        // partial updates up to this one are in the lines
        uint32_t n_lines = 0, n_staging = 0, n_long = 0;
    } base;
    uint64_t frame_key = 0, frame_seq = 0;  // of the frame being recorded
    bool diced_all = false;                 // this frame dices every segment (it becomes the base when it completes)
    uint32_t diced_segments = 0;
    size_t off_dirty = 0, off_ranges = 0, off_counters = 0;  // incremental extras inside host_meta / dev_meta
    uint32_t n_ranges = 0;
    cudaEvent_t fill_done = nullptr;  // recorded on the aux stream after this batch's fill kernel
};

enum CmdKind { CMD_PREPARE, CMD_DRAW };
struct Cmd {
    CmdKind kind;
    int slot;
    int target_page, color_page, clear;
    uint32_t sampling_flags;
    float clear_color[4];
};

}  // namespace

struct pfcu_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    // second stream of the frame: fill scatter + fill run beside propagate + list building (they only meet again in
    // the tile kernel), forked and joined with events so that the same code captures into a graph with two branches
    cudaStream_t aux_stream = nullptr;
    std::vector<cudaEvent_t> sync_events;
    size_t sync_used = 0;
    cudaEvent_t aux_pending = nullptr;  // last event recorded on the aux stream that the main stream has not waited for
    // PFCU_OPT_CONCURRENT_BATCHES: the batches of a frame prepare side by side, batch i on lane i % N_LANES (a lane = the
    // two streams of a batch's kernel chains); only the tile passes stay in order on the context's stream
    static constexpr int N_LANES = 4;
    cudaStream_t lane_main[N_LANES] = {}, lane_aux[N_LANES] = {};
    bool concurrent_batches = true;
    cudaEvent_t frame_fork = nullptr;  // recorded on the context's stream at the frame's first batch: what the lanes wait for
    bool lanes_unjoined = false;       // a lane holds work the context's stream has not waited for (a frame that was abandoned)
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    // asynchronous read-back of the target (pfcu_read_target_async): its own stream, ordered after the frame by an event
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_frame_done = nullptr, ev_copy_done = nullptr;
    bool read_pending = false;  // a D2H copy of the target may still be running: the next frame's writes wait for it
    uint8_t *read_host = nullptr;  // ... into this buffer (re-issued if the frame it was enqueued behind is replayed)
    size_t read_pitch = 0;
    // stroke-to-fill (pfcu_stroke_to_fill): inputs, counts / offsets, result
    DevBuf stroke_in, stroke_counts, stroke_out, stroke_slots, stroke_leaves;
    PinnedBuf stroke_stage;
    std::vector<uint32_t> stroke_host_counts;
    std::vector<uint8_t> stroke_closed;
    // dashing (pfcu_dash_outlines): result layout
    uint32_t dash_n_outlines = 0, dash_points = 0, dash_contours = 0;
    size_t dash_off_flags = 0, dash_off_contours = 0, dash_off_outlines = 0;
    uint32_t stroke_n_contours = 0, stroke_total = 0;
    size_t stroke_off_flags = 0;
    float stroke_ms = 0.f;
    // static resources
    DevBuf lut;
    int lut_w = 0, lut_h = 0;
    cudaArray_t lut_array = nullptr;
    cudaTextureObject_t lut_tex = 0;
    int lut_band = 0;
    DevBuf dummy_px;
    // target
    DevBuf own_target;
    TargetView target{};
    float view_box[4] = {0, 0, 0, 0};
    int origin_tx = 0, origin_ty = 0;
    // scene
    DevBuf scene_dev[2];  // points, then (16-byte aligned) indices: one H2D copy per upload
    size_t indices_off[2] = {0, 0};
    bool stage_busy[2] = {false, false};  // an H2D copy out of stage_scene[i] may still be running
    bool frame_pending = false;           // pfcu_submit_frame without its pfcu_wait_frame yet
    bool pending_served_by_graph = false;
    uint64_t pending_sig = 0;
    uint32_t n_points[2] = {0, 0}, n_segments[2] = {0, 0};
    PinnedBuf stage_scene[2];
    DevBuf paints;  // Paint table decoded from the RGBA16F metadata texels
    uint32_t n_paints = 0;
    int all_solid = 1, unit_range = 1;
    PinnedBuf stage_metadata;
    Page pages[MAX_PAGES];
    PinnedBuf stage_page;
    // frame
    std::vector<BatchSlot> slots;
    int slots_used = 0;
    std::vector<Cmd> cmds;
    bool frame_open = false, in_flight = false, event_begin_recorded = false;
    DevBuf counters;  // BatchCounters[MAX_SLOTS] + frame alpha counter at the end
    DevBuf masks;
    uint32_t mask_cap = 0;
    PinnedBuf host_counters;
    uint32_t launches = 0, retries = 0;
    pfcu_frame_stats last_stats{};
    // per-stage profiling (pfcu_set_profiling): event pairs around every kernel of the frame
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    std::vector<int> prof_stage;  // stage of the kernel that ends at event i + 1 (-1: frame start marker)
    size_t prof_used = 0;
    float stage_ms[PFCU_NUM_STAGES] = {};
    // Retained frame graph: when two consecutive frames enqueue byte-identical work (same kernels, same parameter
    // blocks, same copies -- the signature below), the frame is captured once and later identical frames are ONE graph
    // launch at pfcu_end_frame instead of a dozen kernel launches and as many event operations. While a retained graph
    // exists, pfcu_prepare_batch / pfcu_draw_batch only record; a frame that turns out different is enqueued the
    // ordinary way at pfcu_end_frame and the graph is dropped.
    bool auto_graph = true;
    bool fill_culled_tiles = false;  // PFCU_OPT_FILL_CULLED_TILES
    bool fused_fill = false;         // PFCU_OPT_FUSED_FILL
    int order_groups = 4;            // PFCU_OPT_ORDER_TILE_GROUPS: masked tiles that make a group of 16 tiles expensive
    DevBuf timeline;                 // pfcu_set_timeline: 64-byte header (word 0 = records claimed) + records
    uint32_t timeline_cap = 0;       // 0: off
    // PFCU_OPT_INCREMENTAL_DICE: partial scene updates since the last full upload, per path source
    bool incremental = false;
    struct DirtyRange {
        uint32_t lo, hi;  // global segments [lo, hi)
        uint64_t seq;
    };
    std::vector<DirtyRange> dirty[2];
    uint64_t scene_gen[2] = {1, 1}, update_seq = 0;
    uint32_t uploaded_bytes = 0;  // H2D bytes of the frame being recorded (segments + metadata)
    uint64_t frame_sig = 0, prev_sig = 0, retained_sig = 0;
    cudaGraph_t retained_graph = nullptr;
    cudaGraphExec_t retained_exec = nullptr;
    uint32_t retained_launches = 0;
    bool dry = false;  // this frame is only being recorded (a retained graph may serve it)
    // whole-frame CUDA graph (pfcu_graph_capture)
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    uint32_t graph_launches_per_frame = 0;
    bool capturing = false;
};

namespace {

uint32_t *frame_alpha_counter(pfcu_ctx *c) {
    return reinterpret_cast<uint32_t *>(c->counters.as<BatchCounters>() + MAX_SLOTS);
}

int find_slot(pfcu_ctx *c, uint32_t batch_id) {
    for (int i = 0; i < c->slots_used; i++)
        if (c->slots[i].prepared && c->slots[i].desc.batch_id == batch_id) return i;
    return -1;
}

int sync_if_in_flight(pfcu_ctx *c) {
    if (c->in_flight) {
        if (c->aux_pending) CUDA_TRY(cudaStreamSynchronize(c->aux_stream));
        if (c->lanes_unjoined) {
            for (int i = 0; i < pfcu_ctx::N_LANES; i++) {
                CUDA_TRY(cudaStreamSynchronize(c->lane_main[i]));
                CUDA_TRY(cudaStreamSynchronize(c->lane_aux[i]));
            }
            c->lanes_unjoined = false;
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->in_flight = false;
        c->stage_busy[0] = c->stage_busy[1] = false;
    }
    return PFCU_OK;
}

// Profiling: mark(-1) opens a segment, mark(stage) closes the segment that just ran `stage`.
int prof_mark(pfcu_ctx *c, int stage) {
    if (!c->profiling || c->capturing) return PFCU_OK;
    if (c->prof_used == c->prof_events.size()) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreate(&e));
        c->prof_events.push_back(e);
        c->prof_stage.push_back(-1);
    }
    c->prof_stage[c->prof_used] = stage;
    CUDA_TRY(cudaEventRecord(c->prof_events[c->prof_used], c->stream));
    c->prof_used++;
    return PFCU_OK;
}

// Fork / join between the frame's two streams. Events are reused frame after frame.
int next_sync_event(pfcu_ctx *c, cudaEvent_t *out) {
    if (c->sync_used == c->sync_events.size()) {
        cudaEvent_t e;
        CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->sync_events.push_back(e);
    }
    *out = c->sync_events[c->sync_used++];
    return PFCU_OK;
}

int order_after(pfcu_ctx *c, cudaStream_t later, cudaStream_t earlier, cudaEvent_t *recorded = nullptr) {
    cudaEvent_t e;
    int r = next_sync_event(c, &e);
    if (r) return r;
    CUDA_TRY(cudaEventRecord(e, earlier));
    if (later) CUDA_TRY(cudaStreamWaitEvent(later, e, 0));
    if (recorded) *recorded = e;
    return PFCU_OK;
}

int join_aux(pfcu_ctx *c) {
    if (c->aux_pending) {
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->aux_pending, 0));
        c->aux_pending = nullptr;
    }
    for (int i = 0; i < c->slots_used; i++)  // concurrent batches that nobody drew (clip batches)
        for (int k = 0; k < 2; k++)
            if (!c->slots[i].lane_joined[k] && c->slots[i].lane_done[k]) {
                CUDA_TRY(cudaStreamWaitEvent(c->stream, c->slots[i].lane_done[k], 0));
                c->slots[i].lane_joined[k] = true;
            }
    c->lanes_unjoined = false;  // every lane's work is ordered before what follows on the context's stream
    return PFCU_OK;
}

const char *const STAGE_NAMES[PFCU_NUM_STAGES] = {"pfcu:init (bound)", "pfcu:dice", "pfcu:bin", "pfcu:scan tiles",
                                                 "pfcu:fill scatter", "pfcu:propagate", "pfcu:scan fb", "pfcu:list scatter (sort)",
                                                 "pfcu:fill", "pfcu:tile (composite)"};

// NVTX range over a scope (Nsight Systems / ncu --nvtx show the API call or stage around the kernels it enqueues)
struct NvtxScope {
    explicit NvtxScope(const char *name) { nvtxRangePushA(name); }
    ~NvtxScope() { nvtxRangePop(); }
};

#define LAUNCH_STAGE(stage, expr)        \
    do {                                 \
        NvtxScope _nvtx(STAGE_NAMES[stage]); \
        CUDA_TRY(expr);                  \
        int _r = prof_mark(c, stage);    \
        if (_r) return _r;               \
    } while (0)

void sig_mix(pfcu_ctx *c, const void *data, size_t n) {  // FNV-1a
    const unsigned char *p = static_cast<const unsigned char *>(data);
    uint64_t h = c->frame_sig ? c->frame_sig : 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) h = (h ^ p[i]) * 1099511628211ull;
    c->frame_sig = h;
}

void drop_retained(pfcu_ctx *c) {
    if (c->retained_exec) cudaGraphExecDestroy(c->retained_exec);
    if (c->retained_graph) cudaGraphDestroy(c->retained_graph);
    c->retained_exec = nullptr;
    c->retained_graph = nullptr;
    c->retained_sig = 0;
}

// (Re)build the device view of a slot and enqueue prepare_tiles for it (c->dry: build + sign only).
int enqueue_prepare(pfcu_ctx *c, int slot_index, bool upload_meta = true) {
    BatchSlot &s = c->slots[slot_index];
    const pfcu_batch_desc &d = s.desc;
    const uint32_t fbt = (uint32_t)(((c->target.width + TILE - 1) / TILE) * ((c->target.height + TILE - 1) / TILE));
    const size_t D = std::max<uint32_t>(d.tile_count, 1), C = std::max<uint32_t>(d.column_count, 1);
    const size_t T = std::max<uint32_t>(fbt, 1);
    CUDA_TRY(s.dev_meta.ensure(s.meta_bytes));
    CUDA_TRY(s.tile_word.ensure(D * 4));
    CUDA_TRY(s.fill_begin.ensure(D * 4));
    CUDA_TRY(s.fill_cursor.ensure(D * 4));
    CUDA_TRY(s.col_backdrop.ensure(C * 4));
    CUDA_TRY(s.tile_state.ensure(D * sizeof(TileState)));
    CUDA_TRY(s.lines.ensure((size_t)s.line_cap * sizeof(float4)));
    CUDA_TRY(s.line_meta.ensure((size_t)s.line_cap * sizeof(uint2)));
    CUDA_TRY(s.long_lines.ensure((size_t)s.line_cap * 4));
    CUDA_TRY(s.staging.ensure((size_t)s.staging_cap * sizeof(StagedFill)));
    CUDA_TRY(s.fills.ensure((size_t)s.fill_cap * sizeof(uint2)));
    CUDA_TRY(s.alpha_rank.ensure(D * 4));
    CUDA_TRY(s.fb.ensure(T * sizeof(FbTile)));
    CUDA_TRY(s.prims.ensure(D * sizeof(TilePrim)));
    CUDA_TRY(s.alpha_tiles.ensure(D * sizeof(AlphaTile)));
    CUDA_TRY(s.alpha_map.ensure(D * 4));
    {
        const size_t groups = (T + GROUP_TILES - 1) / GROUP_TILES;
        CUDA_TRY(s.group_of.ensure(groups * 4));
        CUDA_TRY(s.slot_of.ensure(groups * 4));
        CUDA_TRY(s.fb_sorted.ensure(groups * GROUP_TILES * sizeof(FbTile)));
    }
    CUDA_TRY(s.scan_desc0.ensure((D / 2048 + 2) * 8));
    CUDA_TRY(s.scan_desc1.ensure((T / 2048 + 2) * 8));
    BatchView v;
    memset(&v, 0, sizeof(v));  // (padding included: the view is part of the frame signature)
    const char *m = s.dev_meta.as<char>();
    v.backdrops = reinterpret_cast<const pfcu_backdrop_info *>(m + s.off_backdrops);
    v.meta = reinterpret_cast<const pfcu_propagate_metadata *>(m + s.off_meta);
    v.dice = reinterpret_cast<const pfcu_dice_metadata *>(m + s.off_dice);
    v.tpi = reinterpret_cast<const pfcu_tile_path_info *>(m + s.off_tpi);
    const int which = d.path_source ? 1 : 0;
    v.points = c->scene_dev[which].as<float2>();
    v.indices = reinterpret_cast<uint2 *>(c->scene_dev[which].as<char>() + c->indices_off[which]);
    v.n_points = c->n_points[which];
    v.n_segments_total = c->n_segments[which];
    v.path_count = d.path_count;
    v.tile_count = d.tile_count;
    v.segment_count = d.segment_count;
    v.column_count = d.column_count;
    memcpy(v.transform, d.transform, sizeof(v.transform));
    v.identity_transform = d.transform[0] == 1.0f && d.transform[1] == 0.0f && d.transform[2] == 0.0f &&
                           d.transform[3] == 1.0f && d.transform[4] == 0.0f && d.transform[5] == 0.0f;
    memcpy(v.view_box, c->view_box, sizeof(v.view_box));
    v.fb_tw = (c->target.width + TILE - 1) / TILE;
    v.fb_th = (c->target.height + TILE - 1) / TILE;
    v.fb_tx0 = c->origin_tx;
    v.fb_ty0 = c->origin_ty;
    v.counters = c->counters.as<BatchCounters>() + slot_index;
    v.tile_word = s.tile_word.as<uint32_t>();
    v.fill_begin = s.fill_begin.as<uint32_t>();
    v.fill_cursor = s.fill_cursor.as<uint32_t>();
    v.col_backdrop = s.col_backdrop.as<int32_t>();
    v.tile_state = s.tile_state.as<TileState>();
    v.lines = s.lines.as<float4>();
    v.line_meta = s.line_meta.as<uint2>();
    v.long_lines = s.long_lines.as<uint32_t>();
    v.line_capacity = s.line_cap;
    v.staging = s.staging.as<StagedFill>();
    v.staging_capacity = s.staging_cap;
    v.fills = s.fills.as<uint2>();
    v.fill_capacity = s.fill_cap;
    v.alpha_rank = s.alpha_rank.as<uint32_t>();
    v.fb = s.fb.as<FbTile>();
    v.prims = s.prims.as<TilePrim>();
    v.prim_capacity = d.tile_count;
    v.alpha_tiles = s.alpha_tiles.as<AlphaTile>();
    v.alpha_map = s.alpha_map.as<uint32_t>();
    v.cull_fill = d.path_source == 0 && !c->fill_culled_tiles;
    // draw batches only: a clip batch's masks are read by OTHER batches (fill's min() and solid-draw x alpha-clip tiles)
    v.fused_fill = c->fused_fill && d.path_source == 0 && !c->fill_culled_tiles;
    v.group_of = v.slot_of = nullptr;
    v.fb_sorted = nullptr;
    v.group_cost_min = (uint32_t)c->order_groups;
    // (clip batches are never drawn; a framebuffer tile's list has at most one entry per path, and its count shares a
    // word with the count of masked entries)
    if (c->order_groups > 0 && d.path_source == 0 && d.path_count < (1u << FB_COUNT_BITS)) {
        v.group_of = s.group_of.as<uint32_t>();
        v.slot_of = s.slot_of.as<uint32_t>();
        v.fb_sorted = s.fb_sorted.as<FbTile>();
    }
    v.alpha_capacity = d.tile_count;
    v.scan_desc[0] = s.scan_desc0.as<unsigned long long>();
    v.scan_desc[1] = s.scan_desc1.as<unsigned long long>();
    v.clip_meta = nullptr;
    v.clip_tile_state = nullptr;
    v.clip_path_count = 0;
    if (d.clip_batch_id >= 0) {
        const int cs = find_slot(c, (uint32_t)d.clip_batch_id);
        if (cs >= 0 && cs != slot_index) {
            v.clip_meta = c->slots[cs].view.meta;
            v.clip_tile_state = c->slots[cs].view.tile_state;
            v.clip_path_count = c->slots[cs].desc.path_count;
        }
    }
    v.frame_alpha_counter = frame_alpha_counter(c);
    v.masks = c->masks.as<uint8_t>();
    v.mask_capacity = c->mask_cap;
    v.paints = c->paints.as<Paint>();
    v.n_paints = c->n_paints;
    v.solid_prims = c->all_solid;
    v.timeline = c->timeline_cap ? c->timeline.as<uint32_t>() : nullptr;
    v.timeline_capacity = c->timeline_cap;
    if (c->incremental && !s.diced_all) {
        v.dice_ranges = reinterpret_cast<const uint2 *>(m + s.off_ranges);
        v.n_dice_ranges = s.n_ranges;
        v.n_dice_segments = s.diced_segments;
        v.dirty_paths = reinterpret_cast<const uint32_t *>(m + s.off_dirty);
        v.n_static_lines = s.base.n_lines;
    }
    s.view = v;
    s.prepared = true;
    s.lane_done[0] = s.lane_done[1] = nullptr;
    s.lane_joined[0] = s.lane_joined[1] = true;

    PaintView pv;
    memset(&pv, 0, sizeof(pv));
    pv.area_lut = c->lut.as<uint8_t>();
    pv.lut_w = c->lut_w;
    pv.lut_h = c->lut_h;
    pv.lut_tex = c->lut_tex;
    pv.lut_band = c->lut_band;
    {
        const void *host = s.host_meta.p;
        const size_t bytes = s.meta_bytes;
        sig_mix(c, &v, sizeof(v));
        sig_mix(c, &pv, sizeof(pv));
        sig_mix(c, &host, sizeof(host));
        sig_mix(c, &bytes, sizeof(bytes));
    }
    if (c->dry) return PFCU_OK;
    // The frame forks and joins between two streams (graph branches when captured); per-stage profiling keeps
    // everything on one stream so that the event pairs bracket one kernel each.
    //   main: counters = 0, dice ........ bin (short walks) ..... scan, propagate, list building ......... tile
    //   aux :  init (zeroing) ........... bin (long walks) ...... fill scatter ........ fill ............../
    const bool two_streams = !c->profiling || c->capturing;
    // PFCU_OPT_CONCURRENT_BATCHES: this batch's chains run on its lane, beside the other batches' (and beside the tile passes
    // of the batches before it); they start after the frame's first batch was reached on the context's stream (uploads, the
    // zeroed mask slot counter) and after the clip batch this one reads (its tile states and its masks)
    // The FIRST batch of a frame stays on the context's own two streams: a frame of one batch (an SVG) is enqueued exactly
    // as without the option.
    const bool lanes_on = two_streams && c->concurrent_batches;
    const bool lanes = lanes_on && slot_index > 0;
    const int lane = slot_index % pfcu_ctx::N_LANES;
    cudaStream_t ms = lanes ? c->lane_main[lane] : c->stream;
    cudaStream_t aux = lanes ? c->lane_aux[lane] : (two_streams ? c->aux_stream : c->stream);
    if (lanes_on) {
        if (!c->frame_fork) {
            int r = order_after(c, nullptr, c->stream, &c->frame_fork);
            if (r) return r;
        }
        if (lanes) CUDA_TRY(cudaStreamWaitEvent(ms, c->frame_fork, 0));
        if (d.clip_batch_id >= 0) {
            const int cs = find_slot(c, (uint32_t)d.clip_batch_id);
            if (cs >= 0 && cs != slot_index)
                for (int k = 0; k < 2; k++)
                    if (c->slots[cs].lane_done[k]) CUDA_TRY(cudaStreamWaitEvent(ms, c->slots[cs].lane_done[k], 0));
        }
    }
    if (s.meta_bytes && upload_meta) {
        CUDA_TRY(cudaMemcpyAsync(s.dev_meta.p, s.host_meta.p, s.meta_bytes, cudaMemcpyHostToDevice, ms));
        if (!c->capturing) c->uploaded_bytes += (uint32_t)s.meta_bytes;
    }


    {
        int r = prof_mark(c, -1);
        if (r) return r;
    }
    if (two_streams) {
        int r = order_after(c, aux, ms);  // (also orders this batch's aux work after the previous batch's)
        if (r) return r;
    }
    // (one 16 KB memset of every batch's counters at the start of the frame instead of a 64-byte one per batch was
    // measured: 2 us SLOWER per frame -- the small ones do not cost a kernel launch)
    if (v.dice_ranges)  // the retained lines, staging slots and long-line queue stay: the counters start where they ended
        CUDA_TRY(cudaMemcpyAsync(v.counters, m + s.off_counters, sizeof(BatchCounters), cudaMemcpyDeviceToDevice, ms));
    else
        CUDA_TRY(cudaMemsetAsync(v.counters, 0, sizeof(BatchCounters), ms));
    LAUNCH_STAGE(PFCU_STAGE_INIT, launch_init(v, aux));
    LAUNCH_STAGE(PFCU_STAGE_DICE, launch_dice(v, ms));
    if (two_streams) {  // bin needs the zeroed tile words; the long-walk kernel needs dice's lines
        int r = order_after(c, ms, aux);
        if (r) return r;
        r = order_after(c, aux, ms);
        if (r) return r;
    }
    LAUNCH_STAGE(PFCU_STAGE_BIN, launch_bin(v, ms));
    LAUNCH_STAGE(PFCU_STAGE_BIN, launch_bin_long(v, aux));
    if (two_streams) {
        int r = order_after(c, ms, aux);
        if (r) return r;
    }
    LAUNCH_STAGE(PFCU_STAGE_SCAN_TILES, launch_scan_tiles(v, ms));
    // {fill scatter, fill} on the aux stream, {propagate, list building} on the main stream. fill needs propagate's
    // alpha-tile records; the tile kernel (enqueue_draw) needs both branches.
    if (two_streams) {
        int r = order_after(c, aux, ms);
        if (r) return r;
    }
    LAUNCH_STAGE(PFCU_STAGE_FILL_SCATTER, launch_fill_scatter(v, aux));
    LAUNCH_STAGE(PFCU_STAGE_PROPAGATE, launch_propagate(v, ms));
    if (two_streams) {
        int r = order_after(c, aux, ms);
        if (r) return r;
    }
    LAUNCH_STAGE(PFCU_STAGE_SCAN_FB, launch_scan_fb(v, ms));
    LAUNCH_STAGE(PFCU_STAGE_LIST_SCATTER, launch_list_scatter(v, ms));
    LAUNCH_STAGE(PFCU_STAGE_FILL, launch_fill(v, pv, aux));
    s.fill_done = nullptr;
    if (lanes) {  // the ends of both chains: the tile pass (enqueue_draw), batches clipped by this one, the frame's tail
        int r = order_after(c, nullptr, ms, &s.lane_done[0]);
        if (r) return r;
        r = order_after(c, nullptr, aux, &s.lane_done[1]);
        if (r) return r;
        s.lane_joined[0] = s.lane_joined[1] = false;
        c->lanes_unjoined = true;
    } else if (two_streams) {
        int r = order_after(c, nullptr, aux, &s.fill_done);
        if (r) return r;
        c->aux_pending = s.fill_done;
        if (lanes_on) {  // (the frame's first batch: what a batch clipped by it waits for; the context's stream joins as ever)
            r = order_after(c, nullptr, ms, &s.lane_done[0]);
            if (r) return r;
            s.lane_done[1] = s.fill_done;
        }
    }
    c->launches += v.fused_fill ? 9 : 10;
    c->in_flight = true;
    return PFCU_OK;
}

int enqueue_draw(pfcu_ctx *c, const Cmd &cmd) {
    BatchSlot &s = c->slots[cmd.slot];
    TargetView t = c->target;
    int clear = cmd.clear;
    if (cmd.target_page >= 0) {
        if (cmd.target_page >= MAX_PAGES || !c->pages[cmd.target_page].px.p)
            return fail(PFCU_ERR_INVALID, "render target page %d not allocated", cmd.target_page);
        Page &pg = c->pages[cmd.target_page];
        t.pixels = pg.px.as<uint8_t>();
        t.pitch = (size_t)pg.w * 4;
        t.width = pg.w;
        t.height = pg.h;
        clear = 1;  // d3d11/renderer.cpp:382-386
    }
    PaintView pv;
    memset(&pv, 0, sizeof(pv));
    pv.paints = c->paints.as<Paint>();
    pv.n_paints = c->n_paints;
    pv.all_solid = c->all_solid;
    pv.unit_range = c->unit_range;
    pv.color_px = c->dummy_px.as<uint8_t>();
    pv.color_w = pv.color_h = 1;
    pv.sampling_flags = 0;
    if (cmd.color_page >= 0 && cmd.color_page < MAX_PAGES && c->pages[cmd.color_page].px.p) {
        Page &pg = c->pages[cmd.color_page];
        pv.color_px = pg.px.as<uint8_t>();
        pv.color_w = pg.w;
        pv.color_h = pg.h;
        pv.sampling_flags = cmd.sampling_flags;
    }
    pv.area_lut = c->lut.as<uint8_t>();
    pv.lut_w = c->lut_w;
    pv.lut_h = c->lut_h;
    pv.lut_tex = c->lut_tex;
    pv.lut_band = c->lut_band;
    // masks may have been reallocated since the batch was prepared (growth happens only between attempts)
    s.view.masks = c->masks.as<uint8_t>();
    s.view.mask_capacity = c->mask_cap;
    {
        const int origin = cmd.target_page < 0;
        sig_mix(c, &s.view, sizeof(s.view));
        sig_mix(c, &pv, sizeof(pv));
        sig_mix(c, &t.pixels, sizeof(t.pixels));
        sig_mix(c, &t.pitch, sizeof(t.pitch));
        sig_mix(c, &t.width, sizeof(t.width));
        sig_mix(c, &t.height, sizeof(t.height));
        sig_mix(c, &clear, sizeof(clear));
        sig_mix(c, cmd.clear_color, sizeof(cmd.clear_color));
        sig_mix(c, &origin, sizeof(origin));
        const int heavy = s.heavy_paints;
        sig_mix(c, &heavy, sizeof(heavy));
    }
    if (c->dry) return PFCU_OK;
    {
        int r = prof_mark(c, -1);
        if (r) return r;
    }
    for (int k = 0; k < 2; k++)  // concurrent batches: both chains of this batch
        if (!s.lane_joined[k] && s.lane_done[k]) {
            CUDA_TRY(cudaStreamWaitEvent(c->stream, s.lane_done[k], 0));
            s.lane_joined[k] = true;
        }
    if (s.fill_done) {  // join: the masks of this batch (and, the aux stream being in order, of every earlier one)
        CUDA_TRY(cudaStreamWaitEvent(c->stream, s.fill_done, 0));
        if (c->aux_pending == s.fill_done) c->aux_pending = nullptr;
        s.fill_done = nullptr;
    }
    int n_launched = 0;
    LAUNCH_STAGE(PFCU_STAGE_COMPOSITE, launch_composite(s.view, pv, t, clear, cmd.clear_color, cmd.target_page < 0, s.heavy_paints, c->stream, &n_launched));
    c->launches += (uint32_t)n_launched;
    c->in_flight = true;
    return PFCU_OK;
}

// Replays the recorded frame under stream capture. All buffers are already sized by the frame that was just completed,
// so the replay performs no allocation. upload_meta: the graph also copies every batch's metadata from its pinned
// staging buffer (the retained graph, which serves whole frames); otherwise the metadata is taken as resident.
int capture_frame(pfcu_ctx *c, bool upload_meta, cudaGraph_t *graph, cudaGraphExec_t *exec, uint32_t *launches) {
    for (int i = 0; i < c->slots_used; i++) c->slots[i].prepared = false;
    const uint32_t launches_before = c->launches;
    const bool was_dry = c->dry;
    c->dry = false;
    c->capturing = true;
    c->sync_used = 0;
    c->aux_pending = nullptr;
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
        c->capturing = false;
        c->dry = was_dry;
        return fail(PFCU_ERR_CUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(e));
    }
    int rc = PFCU_OK;
    c->frame_fork = nullptr;
    if (cudaMemsetAsync(frame_alpha_counter(c), 0, sizeof(BatchCounters), c->stream) != cudaSuccess) rc = PFCU_ERR_CUDA;
    for (const Cmd &cmd : c->cmds) {
        if (rc) break;
        rc = cmd.kind == CMD_PREPARE ? enqueue_prepare(c, cmd.slot, upload_meta) : enqueue_draw(c, cmd);
    }
    if (!rc) rc = join_aux(c);
    e = cudaStreamEndCapture(c->stream, graph);
    c->capturing = false;
    c->dry = was_dry;
    c->in_flight = false;
    *launches = c->launches - launches_before;
    c->launches = launches_before;
    if (rc) return rc;
    if (e != cudaSuccess) return fail(PFCU_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    CUDA_TRY(cudaGraphInstantiate(exec, *graph, 0));
    return PFCU_OK;
}

}  // namespace

extern "C" {

int pfcu_abi_version(void) { return PFCU_ABI_VERSION; }
const char *pfcu_last_error(void) { return g_last_error.c_str(); }

int pfcu_create(int device_ordinal, pfcu_ctx **out) {
    if (!out) return fail(PFCU_ERR_INVALID, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(PFCU_ERR_CUDA, "no CUDA device (%s); pfcu has no CPU fallback", cudaGetErrorString(e));
    if (device_ordinal < 0 || device_ordinal >= n) return fail(PFCU_ERR_INVALID, "device %d out of range", device_ordinal);
    CUDA_TRY(cudaSetDevice(device_ordinal));
    pfcu_ctx *c = new pfcu_ctx;
    c->device = device_ordinal;
    CUDA_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    for (int i = 0; i < pfcu_ctx::N_LANES; i++) {
        CUDA_TRY(cudaStreamCreateWithFlags(&c->lane_main[i], cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->lane_aux[i], cudaStreamNonBlocking));
    }
    c->stream = c->own_stream;
    CUDA_TRY(cudaEventCreate(&c->ev_begin));
    CUDA_TRY(cudaEventCreate(&c->ev_end));
    CUDA_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_frame_done, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_copy_done, cudaEventDisableTiming));
    CUDA_TRY(c->counters.ensure(sizeof(BatchCounters) * (MAX_SLOTS + 1)));
    CUDA_TRY(cudaMemset(c->counters.p, 0, sizeof(BatchCounters) * (MAX_SLOTS + 1)));
    CUDA_TRY(c->host_counters.ensure(sizeof(BatchCounters) * (MAX_SLOTS + 1)));
    CUDA_TRY(c->dummy_px.ensure(16));
    CUDA_TRY(cudaMemset(c->dummy_px.p, 0, 16));
    c->slots.resize(MAX_SLOTS);
    *out = c;
    return PFCU_OK;
}

void pfcu_destroy(pfcu_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < pfcu_ctx::N_LANES; i++) {
        if (c->lane_main[i]) cudaStreamSynchronize(c->lane_main[i]);
        if (c->lane_aux[i]) cudaStreamSynchronize(c->lane_aux[i]);
    }
    cudaStreamSynchronize(c->aux_stream);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->copy_stream);
    for (auto &s : c->slots) {
        s.host_meta.release();
        for (DevBuf *b : {&s.dev_meta, &s.tile_word, &s.fill_begin, &s.fill_cursor, &s.col_backdrop, &s.tile_state, &s.lines,
                          &s.line_meta, &s.long_lines, &s.staging, &s.fills, &s.fb, &s.alpha_rank, &s.prims,
                          &s.group_of, &s.slot_of, &s.fb_sorted,
                          &s.alpha_tiles, &s.alpha_map, &s.scan_desc0, &s.scan_desc1})
            b->release();
    }
    for (auto &p : c->pages) p.px.release();
    for (int i = 0; i < 2; i++) {
        c->scene_dev[i].release();
        c->stage_scene[i].release();
    }
    c->stroke_in.release();
    c->stroke_counts.release();
    c->stroke_out.release();
    c->stroke_slots.release();
    c->stroke_leaves.release();
    c->stroke_stage.release();
    c->lut.release();
    if (c->lut_tex) cudaDestroyTextureObject(c->lut_tex);
    if (c->lut_array) cudaFreeArray(c->lut_array);
    c->dummy_px.release();
    c->own_target.release();
    c->paints.release();
    c->stage_metadata.release();
    c->stage_page.release();
    c->counters.release();
    c->masks.release();
    c->timeline.release();
    c->host_counters.release();
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    for (cudaEvent_t e : c->prof_events) cudaEventDestroy(e);
    cudaEventDestroy(c->ev_begin);
    cudaEventDestroy(c->ev_end);
    cudaEventDestroy(c->ev_frame_done);
    cudaEventDestroy(c->ev_copy_done);
    cudaStreamDestroy(c->copy_stream);
    drop_retained(c);
    for (cudaEvent_t e : c->sync_events) cudaEventDestroy(e);
    cudaStreamDestroy(c->aux_stream);
    for (int i = 0; i < pfcu_ctx::N_LANES; i++) {
        if (c->lane_main[i]) cudaStreamDestroy(c->lane_main[i]);
        if (c->lane_aux[i]) cudaStreamDestroy(c->lane_aux[i]);
    }
    cudaStreamDestroy(c->own_stream);
    delete c;
}

int pfcu_set_stream(pfcu_ctx *c, void *cuda_stream) {
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    int r = sync_if_in_flight(c);
    if (r) return r;
    c->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : c->own_stream;
    return PFCU_OK;
}

void *pfcu_get_stream(pfcu_ctx *c) { return c ? (void *)c->stream : nullptr; }

int pfcu_set_area_lut(pfcu_ctx *c, const uint8_t *rgba, int width, int height) {
    if (!c || !rgba || width <= 0 || height <= 0) return fail(PFCU_ERR_INVALID, "bad area LUT");
    CUDA_TRY(cudaSetDevice(c->device));
    int r = sync_if_in_flight(c);
    if (r) return r;
    CUDA_TRY(c->lut.ensure((size_t)width * height * 4));
    CUDA_TRY(cudaMemcpy(c->lut.p, rgba, (size_t)width * height * 4, cudaMemcpyHostToDevice));
    c->lut_w = width;
    c->lut_h = height;
    // The fill stage fetches LUT texels through the texture unit (clamp-to-edge, unorm8 -> float) and applies the
    // bilinear weights of fill.comp:70's sampler (core/renderer.cpp:117-165) in fp32.
    if (c->lut_tex) cudaDestroyTextureObject(c->lut_tex);
    if (c->lut_array) cudaFreeArray(c->lut_array);
    c->lut_tex = 0;
    c->lut_array = nullptr;
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    CUDA_TRY(cudaMallocArray(&c->lut_array, &fmt, (size_t)width, (size_t)height));
    CUDA_TRY(cudaMemcpy2DToArray(c->lut_array, 0, 0, rgba, (size_t)width * 4, (size_t)width * 4, (size_t)height,
                                 cudaMemcpyHostToDevice));
    cudaResourceDesc rd{};
    rd.resType = cudaResourceTypeArray;
    rd.res.array.array = c->lut_array;
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 0;
    CUDA_TRY(cudaCreateTextureObject(&c->lut_tex, &rd, &td, nullptr));
    // Saturation band of the LUT (add_group in pfcu_raster.cu): left of 120 - row / 2 every channel must be 255, right of
    // 184 + row / 2 every channel must be 0. True for the reference's area_lut.png; checked, not assumed.
    c->lut_band = width == 256 && height == 256;
    for (int y = 0; y < height && c->lut_band; y++)
        for (int x = 0; x < width; x++) {
            const uint8_t *t = rgba + ((size_t)y * width + x) * 4;
            const bool full = t[0] == 255 && t[1] == 255 && t[2] == 255 && t[3] == 255;
            const bool zero = !t[0] && !t[1] && !t[2] && !t[3];
            if (((float)x < 120.0f - 0.5f * (float)y && !full) || ((float)x > 184.0f + 0.5f * (float)y && !zero)) {
                c->lut_band = 0;
                break;
            }
        }
    return PFCU_OK;
}

int pfcu_set_target(pfcu_ctx *c, int width, int height, void *rgba8_dev, size_t pitch_bytes, const float view_box[4]) {
    if (!c || width <= 0 || height <= 0 || !view_box) return fail(PFCU_ERR_INVALID, "bad target");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->frame_open) return fail(PFCU_ERR_STATE, "the target cannot change inside a frame");
    if (rgba8_dev) {
        if (pitch_bytes < (size_t)width * 4 || (pitch_bytes & 15) || ((uintptr_t)rgba8_dev & 15))
            return fail(PFCU_ERR_INVALID, "target rows must be 16-byte aligned and at least width*4 bytes");
        c->target.pixels = static_cast<uint8_t *>(rgba8_dev);
        c->target.pitch = pitch_bytes;
    } else {
        // The context's own target keeps its contents from frame to frame (LOAD_ACTION_LOAD draws over the previous
        // frame, d3d11/renderer.cpp:382-386); it is zeroed only when it is (re)allocated.
        const size_t pitch = ((size_t)width * 4 + 15) & ~(size_t)15;
        const bool same = c->target.pixels == c->own_target.as<uint8_t>() && c->own_target.p && c->target.width == width &&
                          c->target.height == height && c->target.pitch == pitch;
        if (!same) {
            int r = sync_if_in_flight(c);  // (frames in flight may still write the old buffer)
            if (r) return r;
            CUDA_TRY(c->own_target.ensure(pitch * height));
            CUDA_TRY(cudaMemsetAsync(c->own_target.p, 0, pitch * height, c->stream));
            c->in_flight = true;
        }
        c->target.pixels = c->own_target.as<uint8_t>();
        c->target.pitch = pitch;
    }
    c->target.width = width;
    c->target.height = height;
    memcpy(c->view_box, view_box, sizeof(c->view_box));
    c->origin_tx = c->origin_ty = 0;
    return PFCU_OK;
}

int pfcu_set_target_origin(pfcu_ctx *c, int origin_x, int origin_y) {
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    if ((origin_x % TILE) || (origin_y % TILE)) return fail(PFCU_ERR_INVALID, "the origin must be a multiple of the tile size");
    if (c->frame_open) return fail(PFCU_ERR_STATE, "the origin cannot change inside a frame");
    c->origin_tx = origin_x / TILE;
    c->origin_ty = origin_y / TILE;
    return PFCU_OK;
}

int pfcu_upload_scene(pfcu_ctx *c, int which, const float *points, uint32_t n_points, const uint32_t *indices,
                      uint32_t n_segments) {
    NvtxScope nvtx("pfcu_upload_scene");
    if (!c || which < 0 || which > 1 || (n_points && !points) || (n_segments && !indices))
        return fail(PFCU_ERR_INVALID, "bad scene upload");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->frame_pending) return fail(PFCU_ERR_STATE, "a submitted frame has not been waited for");
    const size_t pb = (size_t)n_points * 8, ib = (size_t)n_segments * 8;
    const size_t ioff = (pb + 15) & ~(size_t)15;
    if (c->stage_busy[which] || ioff + ib > c->scene_dev[which].cap || ioff + ib + 16 > c->stage_scene[which].cap) {
        int r = sync_if_in_flight(c);
        if (r) return r;
    }
    CUDA_TRY(c->scene_dev[which].ensure(std::max<size_t>(ioff + ib, 16)));
    CUDA_TRY(c->stage_scene[which].ensure(ioff + ib + 16));
    char *st = static_cast<char *>(c->stage_scene[which].p);
    if (pb) memcpy(st, points, pb);
    if (ib) memcpy(st + ioff, indices, ib);
    if (pb + ib) CUDA_TRY(cudaMemcpyAsync(c->scene_dev[which].p, st, ioff + ib, cudaMemcpyHostToDevice, c->stream));
    c->indices_off[which] = ioff;
    c->stage_busy[which] = (pb + ib) != 0;
    c->n_points[which] = n_points;
    c->n_segments[which] = n_segments;
    c->in_flight = true;
    c->scene_gen[which]++;  // retained dice output of this source is stale
    c->dirty[which].clear();
    c->uploaded_bytes += (uint32_t)(pb + ib);
    return PFCU_OK;
}

// Scene::epoch / LastSceneInfo::draw_segment_ranges (core/scene.h:32-49, d3d11/scene_builder.cpp:217-218): the points of a
// run of segments changed (a path moved or was reshaped with the same topology); indices and every other segment stay.
int pfcu_update_scene_range(pfcu_ctx *c, int which, uint32_t first_point, const float *points, uint32_t n_points,
                            uint32_t first_segment, uint32_t n_segments) {
    if (!c || which < 0 || which > 1 || !points || !n_points || !n_segments) return fail(PFCU_ERR_INVALID, "bad scene update");
    if (c->frame_pending) return fail(PFCU_ERR_STATE, "a submitted frame has not been waited for");
    if (c->frame_open) return fail(PFCU_ERR_STATE, "the scene cannot change inside a frame");
    if ((uint64_t)first_point + n_points > c->n_points[which] || (uint64_t)first_segment + n_segments > c->n_segments[which])
        return fail(PFCU_ERR_INVALID, "update range exceeds the uploaded scene (%u points, %u segments)", c->n_points[which],
                    c->n_segments[which]);
    CUDA_TRY(cudaSetDevice(c->device));
    NvtxScope nvtx("pfcu_update_scene_range");
    if (c->stage_busy[which]) {  // an earlier copy may still be reading the staging buffer
        int r = sync_if_in_flight(c);
        if (r) return r;
    }
    char *st = static_cast<char *>(c->stage_scene[which].p) + (size_t)first_point * 8;
    memcpy(st, points, (size_t)n_points * 8);  // (the pinned copy stays a mirror of the device arrays)
    CUDA_TRY(cudaMemcpyAsync(c->scene_dev[which].as<char>() + (size_t)first_point * 8, st, (size_t)n_points * 8,
                             cudaMemcpyHostToDevice, c->stream));
    c->stage_busy[which] = true;
    c->in_flight = true;
    c->update_seq++;
    if (c->dirty[which].size() >= 1024) {  // too many to track: fall back to a full re-dice
        c->scene_gen[which]++;
        c->dirty[which].clear();
    } else {
        bool known = false;  // (an animation updates the same path every frame: one entry, its sequence number renewed)
        for (auto &r : c->dirty[which])
            if (r.lo == first_segment && r.hi == first_segment + n_segments) {
                r.seq = c->update_seq;
                known = true;
            }
        if (!known) c->dirty[which].push_back({first_segment, first_segment + n_segments, c->update_seq});
    }
    c->uploaded_bytes += n_points * 8;
    return PFCU_OK;
}

static float half_bits_to_float(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1fu;
    uint32_t man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {  // subnormal half
            int e = -1;
            do {
                e++;
                man <<= 1;
            } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7f800000u | (man << 13);
    } else {
        bits = sign | ((exp + 112u) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

int pfcu_upload_paint_metadata(pfcu_ctx *c, const uint16_t *half_texels, uint32_t n_rows) {
    NvtxScope nvtx("pfcu_upload_paint_metadata");
    if (!c || (n_rows && !half_texels)) return fail(PFCU_ERR_INVALID, "bad metadata upload");
    CUDA_TRY(cudaSetDevice(c->device));
    int r = sync_if_in_flight(c);
    if (r) return r;
    // The reference samples 9 RGBA16F texels per pixel row and layer (tile.comp:707-717); the values only depend on
    // the paint, so they are decoded here once per upload into a float table (exact: half -> float is lossless).
    const uint32_t n_paints = n_rows * 128;  // TEXTURE_METADATA_ENTRIES_PER_ROW
    const size_t bytes = (size_t)std::max<uint32_t>(n_paints, 1) * sizeof(Paint);
    CUDA_TRY(c->paints.ensure(bytes));
    CUDA_TRY(c->stage_metadata.ensure(bytes));
    Paint *table = static_cast<Paint *>(c->stage_metadata.p);
    int all_solid = 1, unit_range = 1;
    for (uint32_t i = 0; i < n_paints; i++) {
        const uint16_t *t = half_texels + ((size_t)(i / 128) * 1280 + (size_t)(i % 128) * 10) * 4;
        auto texel = [&](int e) {
            return make_float4(half_bits_to_float(t[e * 4 + 0]), half_bits_to_float(t[e * 4 + 1]),
                               half_bits_to_float(t[e * 4 + 2]), half_bits_to_float(t[e * 4 + 3]));
        };
        Paint p{};
        p.m0 = texel(0);
        p.m1 = texel(1);
        p.base = texel(2);
        p.fp0 = texel(3);
        p.fp1 = texel(4);
        p.fp2 = texel(5);
        p.fp3 = texel(6);
        p.fp4 = texel(7);
        p.ctrl = (int32_t)texel(8).x;  // int(extra.x), tile.comp:725
        if (p.ctrl != 0) all_solid = 0;
        // filters: none, radial gradient (0x1), text (0x2), blur (0x3), colour matrix (0x4) -- tile.comp:407-456. Upstream
        // emits neither text nor colour matrix (paint/palette.cpp:65-67) and binds a 1 x 1 dummy as the text filter's gamma
        // LUT (d3d11/renderer.cpp:262-266): a paint that turns gamma correction on has no defined result and is refused
        const int filter = (p.ctrl >> 4) & 0xf;
        if (((p.ctrl >> 8) & 0x3) != 0 && filter > 0x4)
            return fail(PFCU_ERR_INVALID, "paint %u uses the unknown filter %d", i, filter);
        if (((p.ctrl >> 8) & 0x3) != 0 && filter == 0x2 && p.fp2.w != 0.0f)
            return fail(PFCU_ERR_INVALID, "paint %u: the text filter's gamma correction needs a gamma LUT the reference does not upload", i);
        for (float v : {p.base.x, p.base.y, p.base.z, p.base.w})
            if (!(v >= 0.0f && v <= 1.0f)) unit_range = 0;
        table[i] = p;
    }
    if (n_paints) {
        CUDA_TRY(cudaMemcpyAsync(c->paints.p, table, (size_t)n_paints * sizeof(Paint), cudaMemcpyHostToDevice, c->stream));
        c->in_flight = true;
    }
    c->n_paints = n_paints;
    c->all_solid = all_solid;
    c->unit_range = unit_range;
    return PFCU_OK;
}

int pfcu_alloc_page(pfcu_ctx *c, uint32_t page, int width, int height) {
    if (!c || page >= MAX_PAGES || width <= 0 || height <= 0) return fail(PFCU_ERR_INVALID, "bad page");
    CUDA_TRY(cudaSetDevice(c->device));
    int r = sync_if_in_flight(c);
    if (r) return r;
    Page &pg = c->pages[page];
    CUDA_TRY(pg.px.ensure((size_t)width * height * 4));
    CUDA_TRY(cudaMemset(pg.px.p, 0, (size_t)width * height * 4));
    pg.w = width;
    pg.h = height;
    return PFCU_OK;
}

int pfcu_upload_page_region(pfcu_ctx *c, uint32_t page, int x, int y, int width, int height, const uint8_t *rgba) {
    if (!c || page >= MAX_PAGES || !c->pages[page].px.p) return fail(PFCU_ERR_INVALID, "texture page not allocated");
    Page &pg = c->pages[page];
    if (!rgba || x < 0 || y < 0 || width <= 0 || height <= 0 || x + width > pg.w || y + height > pg.h)
        return fail(PFCU_ERR_INVALID, "tried to write invalid region of a texture");
    CUDA_TRY(cudaSetDevice(c->device));
    int r = sync_if_in_flight(c);
    if (r) return r;
    CUDA_TRY(cudaMemcpy2D(pg.px.as<uint8_t>() + ((size_t)y * pg.w + x) * 4, (size_t)pg.w * 4, rgba, (size_t)width * 4,
                          (size_t)width * 4, height, cudaMemcpyHostToDevice));
    return PFCU_OK;
}

// ------------------------------------------------------------------------------------------------ stroke-to-fill

int pfcu_stroke_to_fill(pfcu_ctx *c, const float *points, const uint8_t *flags, uint32_t n_points, const uint32_t *contour_first,
                        const uint8_t *closed, const uint32_t *style_index, uint32_t n_contours, const pfcu_stroke_style *styles,
                        uint32_t n_styles, uint32_t *n_out_contours, uint32_t *n_out_points) {
    if (!c || (n_points && (!points || !flags)) || (n_contours && (!contour_first || !closed || !style_index || !styles || !n_styles)))
        return fail(PFCU_ERR_INVALID, "bad stroke input");
    // segment slots per contour: a segment starts at an on-curve point, plus the closing line; two sides
    std::vector<uint32_t> seg_first((size_t)n_contours + 1, 0u);
    for (uint32_t i = 0; i < n_contours; i++) {
        if (contour_first[i] > contour_first[i + 1] || contour_first[i + 1] > n_points)
            return fail(PFCU_ERR_INVALID, "contour %u: point range [%u, %u) outside the %u points", i, contour_first[i], contour_first[i + 1], n_points);
        if (style_index[i] >= n_styles) return fail(PFCU_ERR_INVALID, "contour %u: style %u of %u", i, style_index[i], n_styles);
        uint32_t on_curve = 0;
        for (uint32_t k = contour_first[i]; k < contour_first[i + 1]; k++) on_curve += flags[k] == 0;
        seg_first[i + 1] = seg_first[i] + 2u * (on_curve + 1u);
    }
    const uint32_t n_slots = n_contours ? seg_first[n_contours] : 0u;
    CUDA_TRY(cudaSetDevice(c->device));
    NvtxScope nvtx("pfcu_stroke_to_fill");
    // one staged H2D copy: points | flags | contour_first | closed | style_index | styles | seg_first (each 16-byte aligned)
    auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_pts = 0, o_flags = align16(o_pts + (size_t)n_points * 8), o_first = align16(o_flags + n_points);
    const size_t o_closed = align16(o_first + ((size_t)n_contours + 1) * 4), o_style = align16(o_closed + n_contours);
    const size_t o_styles = align16(o_style + (size_t)n_contours * 4), o_seg = align16(o_styles + (size_t)n_styles * sizeof(pfcu_stroke_style));
    const size_t bytes = align16(o_seg + ((size_t)n_contours + 1) * 4);
    CUDA_TRY(cudaStreamSynchronize(c->stream));  // (the staging buffer of an earlier call may still be in flight)
    CUDA_TRY(c->stroke_stage.ensure(bytes + 16));
    CUDA_TRY(c->stroke_in.ensure(bytes + 16));
    const size_t n_scan = std::max<size_t>(n_slots, 2 * (size_t)n_contours);
    // counts | offsets (2 per contour) | leaf_count | leaf_offset (per slot) | total | scan scratch
    const size_t w_counts = 0, w_offsets = 2 * (size_t)n_contours, w_lcount = 4 * (size_t)n_contours, w_loffset = w_lcount + n_slots;
    const size_t w_total = w_loffset + n_slots, w_scratch = w_total + 4, w_end = w_scratch + n_scan / 4096 + 2;
    CUDA_TRY(c->stroke_counts.ensure(w_end * 4));
    CUDA_TRY(c->stroke_slots.ensure(std::max<size_t>((size_t)n_slots * stroke_slot_bytes(), 16)));
    char *st = static_cast<char *>(c->stroke_stage.p);
    if (n_points) {
        memcpy(st + o_pts, points, (size_t)n_points * 8);
        memcpy(st + o_flags, flags, n_points);
    }
    if (n_contours) {
        memcpy(st + o_first, contour_first, ((size_t)n_contours + 1) * 4);
        memcpy(st + o_closed, closed, n_contours);
        memcpy(st + o_style, style_index, (size_t)n_contours * 4);
        memcpy(st + o_styles, styles, (size_t)n_styles * sizeof(pfcu_stroke_style));
        memcpy(st + o_seg, seg_first.data(), ((size_t)n_contours + 1) * 4);
    }
    CUDA_TRY(cudaMemcpyAsync(c->stroke_in.p, st, bytes, cudaMemcpyHostToDevice, c->stream));
    const char *d = c->stroke_in.as<char>();
    uint32_t *w = c->stroke_counts.as<uint32_t>();
    StrokeArgs a{};
    a.pts = reinterpret_cast<const float2 *>(d + o_pts);
    a.flags = reinterpret_cast<const uint8_t *>(d + o_flags);
    a.contour_first = reinterpret_cast<const uint32_t *>(d + o_first);
    a.closed = reinterpret_cast<const uint8_t *>(d + o_closed);
    a.style_index = reinterpret_cast<const uint32_t *>(d + o_style);
    a.styles = reinterpret_cast<const pfcu_stroke_style *>(d + o_styles);
    a.seg_first = reinterpret_cast<const uint32_t *>(d + o_seg);
    a.n_contours = n_contours;
    a.n_slots = n_slots;
    a.slots = c->stroke_slots.p;
    a.counts = w + w_counts;
    a.offsets = w + w_offsets;
    a.leaf_count = w + w_lcount;
    a.leaf_offset = w + w_loffset;
    a.total = w + w_total;
    a.scratch = w + w_scratch;
    c->stroke_n_contours = n_contours;
    c->stroke_total = 0;
    c->stroke_host_counts.assign(2 * (size_t)n_contours, 0u);
    c->stroke_closed.assign(closed, closed + n_contours);
    if (n_contours) {
        CUDA_TRY(cudaEventRecord(c->ev_begin, c->stream));
        // level 1 + 2a: segments of every contour side, leaves per segment counted and scanned
        CUDA_TRY(launch_stroke_segments(a, c->stream));
        uint32_t total_leaves = 0;
        CUDA_TRY(cudaMemcpyAsync(&total_leaves, a.total, 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // read-back 1: the leaf total sizes the leaf array
        CUDA_TRY(c->stroke_leaves.ensure(std::max<size_t>((size_t)total_leaves * stroke_leaf_bytes(), 16)));
        a.leaves = c->stroke_leaves.p;
        // level 2b + 3a: leaves written, points per output contour counted and scanned
        CUDA_TRY(launch_stroke_count(a, c->stream));
        CUDA_TRY(cudaMemcpyAsync(c->stroke_host_counts.data(), a.counts, 2 * (size_t)n_contours * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(&c->stroke_total, a.total, 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));  // read-back 2: the point total sizes the output
        c->stroke_off_flags = align16((size_t)c->stroke_total * 8);
        CUDA_TRY(c->stroke_out.ensure(c->stroke_off_flags + c->stroke_total + 16));
        // level 3b: the output contours
        CUDA_TRY(launch_stroke_write(a, c->stroke_out.as<float2>(), c->stroke_out.as<uint8_t>() + c->stroke_off_flags, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev_end, c->stream));
    }
    uint32_t n_out = 0;
    for (uint32_t i = 0; i < n_contours; i++) n_out += closed[i] ? 2u : 1u;
    if (n_out_contours) *n_out_contours = n_out;
    if (n_out_points) *n_out_points = c->stroke_total;
    c->in_flight = true;
    return PFCU_OK;
}

int pfcu_stroke_result(pfcu_ctx *c, float *out_points, uint8_t *out_flags, uint32_t *out_contour_first) {
    if (!c || !out_contour_first || (c->stroke_total && (!out_points || !out_flags))) return fail(PFCU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->stroke_total) {
        CUDA_TRY(cudaMemcpyAsync(out_points, c->stroke_out.p, (size_t)c->stroke_total * 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(out_flags, c->stroke_out.as<uint8_t>() + c->stroke_off_flags, c->stroke_total, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->stroke_ms = 0.f;
    if (c->stroke_n_contours) cudaEventElapsedTime(&c->stroke_ms, c->ev_begin, c->ev_end);
    // output contours in input order: two per closed input contour (outer, inner), one per open contour
    uint32_t at = 0, k = 0;
    const std::vector<uint32_t> &cnt = c->stroke_host_counts;
    for (uint32_t i = 0; i < c->stroke_n_contours; i++) {
        out_contour_first[k++] = at;
        at += cnt[2 * i];
        if (c->stroke_closed[i]) {
            out_contour_first[k++] = at;
            at += cnt[2 * i + 1];
        }
    }
    out_contour_first[k] = at;
    return PFCU_OK;
}

float pfcu_stroke_gpu_ms(pfcu_ctx *c) { return c ? c->stroke_ms : 0.f; }

int pfcu_dash_outlines(pfcu_ctx *c, const float *points, const uint8_t *flags, uint32_t n_points, const uint32_t *contour_first,
                       const uint8_t *closed, uint32_t n_contours, const uint32_t *outline_first, uint32_t n_outlines,
                       const float *dashes, const uint32_t *dash_first, const float *dash_offset, uint32_t *n_out_contours,
                       uint32_t *n_out_points) {
    if (!c || (n_points && (!points || !flags)) || (n_contours && (!contour_first || !closed)) ||
        (n_outlines && (!outline_first || !dashes || !dash_first || !dash_offset)))
        return fail(PFCU_ERR_INVALID, "bad dash input");
    for (uint32_t i = 0; i < n_contours; i++)
        if (contour_first[i] > contour_first[i + 1] || contour_first[i + 1] > n_points)
            return fail(PFCU_ERR_INVALID, "contour %u: point range outside the %u points", i, n_points);
    for (uint32_t o = 0; o < n_outlines; o++)
        if (outline_first[o] > outline_first[o + 1] || outline_first[o + 1] > n_contours || dash_first[o] > dash_first[o + 1])
            return fail(PFCU_ERR_INVALID, "outline %u: bad contour or dash range", o);
    CUDA_TRY(cudaSetDevice(c->device));
    NvtxScope nvtx("pfcu_dash_outlines");
    const uint32_t n_dash = n_outlines ? dash_first[n_outlines] : 0u;
    auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    const size_t o_pts = 0, o_flags = align16((size_t)n_points * 8), o_first = align16(o_flags + n_points);
    const size_t o_closed = align16(o_first + ((size_t)n_contours + 1) * 4), o_ofirst = align16(o_closed + n_contours);
    const size_t o_dashes = align16(o_ofirst + ((size_t)n_outlines + 1) * 4), o_dfirst = align16(o_dashes + (size_t)n_dash * 4);
    const size_t o_doff = align16(o_dfirst + ((size_t)n_outlines + 1) * 4), bytes = align16(o_doff + (size_t)n_outlines * 4);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(c->stroke_stage.ensure(bytes + 16));
    CUDA_TRY(c->stroke_in.ensure(bytes + 16));
    CUDA_TRY(c->stroke_counts.ensure(((size_t)n_outlines * 4 + 8 + n_outlines / 4096 + 2) * 4));
    char *st = static_cast<char *>(c->stroke_stage.p);
    if (n_points) {
        memcpy(st + o_pts, points, (size_t)n_points * 8);
        memcpy(st + o_flags, flags, n_points);
    }
    if (n_contours) {
        memcpy(st + o_first, contour_first, ((size_t)n_contours + 1) * 4);
        memcpy(st + o_closed, closed, n_contours);
    }
    if (n_outlines) {
        memcpy(st + o_ofirst, outline_first, ((size_t)n_outlines + 1) * 4);
        memcpy(st + o_dashes, dashes, (size_t)n_dash * 4);
        memcpy(st + o_dfirst, dash_first, ((size_t)n_outlines + 1) * 4);
        memcpy(st + o_doff, dash_offset, (size_t)n_outlines * 4);
    }
    CUDA_TRY(cudaMemcpyAsync(c->stroke_in.p, st, bytes, cudaMemcpyHostToDevice, c->stream));
    const char *d = c->stroke_in.as<char>();
    uint32_t *pc = c->stroke_counts.as<uint32_t>(), *cc = pc + n_outlines, *po = cc + n_outlines, *co = po + n_outlines, *totals = co + n_outlines, *scratch = totals + 4;
    uint32_t host_totals[2] = {0, 0};
    c->dash_n_outlines = n_outlines;
    c->dash_points = c->dash_contours = 0;
    if (n_outlines) {
        CUDA_TRY(cudaEventRecord(c->ev_begin, c->stream));
        auto P = [&](size_t off) { return d + off; };
        CUDA_TRY(launch_dash_count(reinterpret_cast<const float2 *>(P(o_pts)), reinterpret_cast<const uint8_t *>(P(o_flags)),
                                   reinterpret_cast<const uint32_t *>(P(o_first)), reinterpret_cast<const uint8_t *>(P(o_closed)),
                                   reinterpret_cast<const uint32_t *>(P(o_ofirst)), reinterpret_cast<const float *>(P(o_dashes)),
                                   reinterpret_cast<const uint32_t *>(P(o_dfirst)), reinterpret_cast<const float *>(P(o_doff)), n_outlines,
                                   pc, cc, po, co, totals, scratch, c->stream));
        CUDA_TRY(cudaMemcpyAsync(host_totals, totals, 8, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->dash_points = host_totals[0];
        c->dash_contours = host_totals[1];
        c->dash_off_flags = align16((size_t)c->dash_points * 8);
        c->dash_off_contours = align16(c->dash_off_flags + c->dash_points);
        c->dash_off_outlines = align16(c->dash_off_contours + ((size_t)c->dash_contours + 1) * 4);
        CUDA_TRY(c->stroke_out.ensure(c->dash_off_outlines + ((size_t)n_outlines + 1) * 4 + 16));
        char *o = c->stroke_out.as<char>();
        CUDA_TRY(cudaMemsetAsync(o + c->dash_off_contours, 0, 4, c->stream));
        CUDA_TRY(launch_dash_write(reinterpret_cast<const float2 *>(P(o_pts)), reinterpret_cast<const uint8_t *>(P(o_flags)),
                                   reinterpret_cast<const uint32_t *>(P(o_first)), reinterpret_cast<const uint8_t *>(P(o_closed)),
                                   reinterpret_cast<const uint32_t *>(P(o_ofirst)), reinterpret_cast<const float *>(P(o_dashes)),
                                   reinterpret_cast<const uint32_t *>(P(o_dfirst)), reinterpret_cast<const float *>(P(o_doff)), n_outlines,
                                   pc, cc, po, co, reinterpret_cast<float2 *>(o), reinterpret_cast<uint8_t *>(o + c->dash_off_flags),
                                   reinterpret_cast<uint32_t *>(o + c->dash_off_contours), c->stream));
        // contour ranges per outline = the exclusive scan of the contour counts (+ the total)
        CUDA_TRY(cudaMemcpyAsync(o + c->dash_off_outlines, co, (size_t)n_outlines * 4, cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaMemcpyAsync(o + c->dash_off_outlines + (size_t)n_outlines * 4, totals + 1, 4, cudaMemcpyDeviceToDevice, c->stream));
        CUDA_TRY(cudaEventRecord(c->ev_end, c->stream));
    }
    if (n_out_contours) *n_out_contours = c->dash_contours;
    if (n_out_points) *n_out_points = c->dash_points;
    c->in_flight = true;
    return PFCU_OK;
}

int pfcu_dash_result(pfcu_ctx *c, float *out_points, uint8_t *out_flags, uint32_t *out_contour_first, uint32_t *out_outline_first) {
    if (!c || !out_contour_first || !out_outline_first || (c->dash_points && (!out_points || !out_flags)))
        return fail(PFCU_ERR_INVALID, "null argument");
    CUDA_TRY(cudaSetDevice(c->device));
    out_contour_first[0] = 0;
    out_outline_first[0] = 0;
    if (c->dash_n_outlines) {
        const char *o = c->stroke_out.as<char>();
        if (c->dash_points) {
            CUDA_TRY(cudaMemcpyAsync(out_points, o, (size_t)c->dash_points * 8, cudaMemcpyDeviceToHost, c->stream));
            CUDA_TRY(cudaMemcpyAsync(out_flags, o + c->dash_off_flags, c->dash_points, cudaMemcpyDeviceToHost, c->stream));
        }
        CUDA_TRY(cudaMemcpyAsync(out_contour_first, o + c->dash_off_contours, ((size_t)c->dash_contours + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaMemcpyAsync(out_outline_first, o + c->dash_off_outlines, ((size_t)c->dash_n_outlines + 1) * 4, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->stroke_ms = 0.f;
    if (c->dash_n_outlines) cudaEventElapsedTime(&c->stroke_ms, c->ev_begin, c->ev_end);
    return PFCU_OK;
}

// ------------------------------------------------------------------------------------------------ frame

int pfcu_begin_frame(pfcu_ctx *c) {
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    if (!c->target.pixels) return fail(PFCU_ERR_STATE, "pfcu_set_target has not been called");
    if (!c->lut_tex) return fail(PFCU_ERR_STATE, "pfcu_set_area_lut has not been called (the fill stage has no coverage table)");
    if (c->frame_pending) return fail(PFCU_ERR_STATE, "a submitted frame has not been waited for");
    CUDA_TRY(cudaSetDevice(c->device));
    if (c->read_pending) {  // the copy engine is still reading the target this frame is about to overwrite
        CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_copy_done, 0));
        c->read_pending = false;
    }
    for (int i = 0; i < c->slots_used; i++) c->slots[i].prepared = false;
    c->slots_used = 0;
    c->cmds.clear();
    c->launches = 0;
    c->retries = 0;
    c->frame_open = true;
    c->event_begin_recorded = false;
    c->prof_used = 0;
    c->sync_used = 0;
    c->aux_pending = nullptr;
    c->frame_sig = 0;
    c->dry = c->auto_graph && !c->profiling && c->retained_exec != nullptr;
    return PFCU_OK;
}

int pfcu_prepare_batch(pfcu_ctx *c, const pfcu_batch_desc *d) {
    NvtxScope nvtx("pfcu_prepare_batch");
    if (!c || !d) return fail(PFCU_ERR_INVALID, "null argument");
    if (!c->frame_open) return fail(PFCU_ERR_STATE, "pfcu_begin_frame has not been called");
    if (c->slots_used >= MAX_SLOTS) return fail(PFCU_ERR_INVALID, "too many batches in one frame");
    if ((d->path_count && (!d->propagate_metadata || !d->dice_metadata || !d->tile_path_info)) ||
        (d->column_count && !d->backdrops))
        return fail(PFCU_ERR_INVALID, "batch metadata missing");
    if (d->tile_count >= (1u << 24)) return fail(PFCU_ERR_INVALID, "tile_count exceeds the 24-bit alpha tile id range");
    if (d->segment_count && !d->path_count) return fail(PFCU_ERR_INVALID, "a batch with segments needs at least one path");
    // the kernels index the dense tile maps and the column backdrops straight from this metadata: check it once, here
    for (uint32_t i = 0; i < d->path_count; i++) {
        const pfcu_propagate_metadata &m = d->propagate_metadata[i];
        const int64_t w = (int64_t)m.tile_rect[2] - m.tile_rect[0], h = (int64_t)m.tile_rect[3] - m.tile_rect[1];
        if (w <= 0 || h <= 0) continue;  // an empty rect owns no tiles
        if ((uint64_t)m.tile_offset + (uint64_t)(w * h) > d->tile_count)
            return fail(PFCU_ERR_INVALID, "path %u: tile_offset %u + %lld x %lld tiles exceeds tile_count %u", i, m.tile_offset,
                        (long long)w, (long long)h, d->tile_count);
        if ((uint64_t)m.backdrop_offset + (uint64_t)w > d->column_count)
            return fail(PFCU_ERR_INVALID, "path %u: backdrop_offset %u + %lld columns exceeds column_count %u", i,
                        m.backdrop_offset, (long long)w, d->column_count);
    }
    for (uint32_t i = 0; i < d->column_count; i++)
        if (d->backdrops[i].path_index >= d->path_count)
            return fail(PFCU_ERR_INVALID, "backdrop column %u names path %u of %u", i, d->backdrops[i].path_index, d->path_count);
    CUDA_TRY(cudaSetDevice(c->device));
    const int slot_index = c->slots_used;
    BatchSlot &s = c->slots[slot_index];
    if (!c->mask_cap) {
        c->mask_cap = 16384;
        CUDA_TRY(c->masks.ensure((size_t)c->mask_cap * 256));
    }
    if (!c->event_begin_recorded && !c->dry) {
        CUDA_TRY(cudaEventRecord(c->ev_begin, c->stream));
        c->event_begin_recorded = true;
        // first batch of the frame: reset the frame-global alpha tile counter
        c->frame_fork = nullptr;
        CUDA_TRY(cudaMemsetAsync(frame_alpha_counter(c), 0, sizeof(BatchCounters), c->stream));
    }
    s.desc = *d;
    // pack the metadata vectors into pinned memory (stable for the async copy and for replay)
    auto align16 = [](size_t v) { return (v + 15) & ~(size_t)15; };
    s.off_backdrops = 0;
    s.off_meta = align16(s.off_backdrops + (size_t)d->column_count * sizeof(pfcu_backdrop_info));
    s.off_dice = align16(s.off_meta + (size_t)d->path_count * sizeof(pfcu_propagate_metadata));
    s.off_tpi = align16(s.off_dice + (size_t)d->path_count * sizeof(pfcu_dice_metadata));
    s.meta_bytes = align16(s.off_tpi + (size_t)d->path_count * sizeof(pfcu_tile_path_info));
    // incremental extras: dirty-path bitmap, dice ranges (+ sentinel), the seed of the batch counters
    s.off_dirty = s.meta_bytes;
    s.off_ranges = align16(s.off_dirty + ((size_t)d->path_count + 31) / 32 * 4);
    s.off_counters = align16(s.off_ranges + ((size_t)d->path_count + 1) * sizeof(uint2));
    if (c->incremental) s.meta_bytes = s.off_counters + sizeof(BatchCounters);
    CUDA_TRY(s.host_meta.ensure(std::max<size_t>(s.meta_bytes, 16)));
    char *hm = static_cast<char *>(s.host_meta.p);
    if (d->column_count) memcpy(hm + s.off_backdrops, d->backdrops, (size_t)d->column_count * sizeof(pfcu_backdrop_info));
    if (d->path_count) {
        memcpy(hm + s.off_meta, d->propagate_metadata, (size_t)d->path_count * sizeof(pfcu_propagate_metadata));
        memcpy(hm + s.off_dice, d->dice_metadata, (size_t)d->path_count * sizeof(pfcu_dice_metadata));
        memcpy(hm + s.off_tpi, d->tile_path_info, (size_t)d->path_count * sizeof(pfcu_tile_path_info));
    }
    s.heavy_paints = false;
    if (!c->all_solid && c->stage_metadata.p) {
        const Paint *table = static_cast<const Paint *>(c->stage_metadata.p);
        for (uint32_t i = 0; i < d->path_count && !s.heavy_paints; i++) {
            const uint32_t paint = d->tile_path_info[i].color;
            if (paint < c->n_paints) {
                const int ctrl = table[paint].ctrl;
                s.heavy_paints = ((ctrl >> 8) & 0x3) != 0 && ((ctrl >> 4) & 0xf) == 0x3;
            }
        }
    }
    // ---- incremental frames: which paths changed since the slot's retained dice output was made
    s.diced_all = true;
    s.diced_segments = d->segment_count;
    s.n_ranges = 0;
    s.frame_seq = c->update_seq;
    if (c->incremental && d->path_count) {
        const int which = d->path_source ? 1 : 0;
        uint64_t key = 1469598103934665603ull;
        auto mix = [&](const void *p, size_t n) {  // (8 bytes per step: the dice metadata of a glyph-density batch is megabytes)
            const unsigned char *q = static_cast<const unsigned char *>(p);
            size_t i = 0;
            for (; i + 8 <= n; i += 8) {
                uint64_t w;
                memcpy(&w, q + i, 8);
                key = (key ^ w) * 1099511628211ull;
                key ^= key >> 29;
            }
            for (; i < n; i++) key = (key ^ q[i]) * 1099511628211ull;
        };
        mix(&d->batch_id, 4); mix(&d->path_count, 4); mix(&d->segment_count, 4); mix(&d->path_source, 4);
        mix(d->transform, sizeof(d->transform)); mix(c->view_box, sizeof(c->view_box));
        mix(d->dice_metadata, (size_t)d->path_count * sizeof(pfcu_dice_metadata));
        s.frame_key = key;
        const bool usable = s.base.valid && s.base.key == key && s.base.scene_gen == c->scene_gen[which];
        if (usable) {
            uint32_t *bitmap = reinterpret_cast<uint32_t *>(hm + s.off_dirty);
            uint2 *ranges = reinterpret_cast<uint2 *>(hm + s.off_ranges);
            memset(bitmap, 0, ((size_t)d->path_count + 31) / 32 * 4);
            uint32_t n_ranges = 0, n_seg = 0;
            for (uint32_t p = 0; p < d->path_count; p++) {
                const pfcu_dice_metadata &dm = d->dice_metadata[p];
                const uint32_t cnt = (p + 1 < d->path_count ? d->dice_metadata[p + 1].first_batch_segment_index : d->segment_count) -
                                     dm.first_batch_segment_index;
                const uint32_t g0 = dm.first_global_segment_index, g1 = g0 + cnt;
                bool dirty = false;
                for (const auto &r : c->dirty[which])
                    if (r.seq > s.base.seq && r.lo < g1 && g0 < r.hi) {
                        dirty = true;
                        break;
                    }
                if (!dirty || !cnt) continue;
                bitmap[p >> 5] |= 1u << (p & 31u);
                if (n_ranges && ranges[n_ranges - 1].x + (n_seg - ranges[n_ranges - 1].y) == dm.first_batch_segment_index) {
                    // (extends the previous range)
                } else {
                    ranges[n_ranges++] = make_uint2(dm.first_batch_segment_index, n_seg);
                }
                n_seg += cnt;
            }
            if (n_seg * 2 <= d->segment_count) {  // otherwise a full re-dice is cheaper, and it renews the base
                ranges[n_ranges] = make_uint2(0xffffffffu, n_seg);
                s.n_ranges = n_ranges;
                s.diced_all = false;
                s.diced_segments = n_seg;
                BatchCounters seed{};
                seed.n_lines = s.base.n_lines;
                seed.n_staging = s.base.n_staging;
                seed.n_long = s.base.n_long;
                memcpy(hm + s.off_counters, &seed, sizeof(seed));
            }
        }
    }
    s.desc.backdrops = nullptr;  // the caller's vectors are not retained
    s.desc.propagate_metadata = nullptr;
    s.desc.dice_metadata = nullptr;
    s.desc.tile_path_info = nullptr;
    // capacities persist across frames; first guess from the segment count (dice.comp's 16K start, renderer.cpp:45)
    s.line_cap = std::max<uint32_t>(s.line_cap, (uint32_t)round_up_pow2(std::max<size_t>(16384, (size_t)d->segment_count * 8)));
    s.fill_cap = std::max<uint32_t>(s.fill_cap, (uint32_t)round_up_pow2(std::max<size_t>(65536, (size_t)s.line_cap * 2)));
    s.staging_cap = std::max<uint32_t>(s.staging_cap, (uint32_t)round_up_pow2((size_t)s.line_cap * 4));
    c->slots_used++;
    Cmd cmd{};
    cmd.kind = CMD_PREPARE;
    cmd.slot = slot_index;
    c->cmds.push_back(cmd);
    return enqueue_prepare(c, slot_index);
}

int pfcu_draw_batch(pfcu_ctx *c, uint32_t batch_id, int target_page, int color_page, uint32_t sampling_flags, int clear,
                    const float clear_color[4]) {
    NvtxScope nvtx("pfcu_draw_batch");
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    if (!c->frame_open) return fail(PFCU_ERR_STATE, "pfcu_begin_frame has not been called");
    const int slot = find_slot(c, batch_id);
    if (slot < 0) return fail(PFCU_ERR_STATE, "batch %u has not been prepared in this frame", batch_id);
    CUDA_TRY(cudaSetDevice(c->device));
    Cmd cmd{};
    cmd.kind = CMD_DRAW;
    cmd.slot = slot;
    cmd.target_page = target_page;
    cmd.color_page = color_page;
    cmd.sampling_flags = sampling_flags;
    cmd.clear = clear;
    for (int i = 0; i < 4; i++) cmd.clear_color[i] = clear_color ? clear_color[i] : 0.0f;
    c->cmds.push_back(cmd);
    return enqueue_draw(c, cmd);
}

// The tail of every attempt: join the second stream, close the frame's event pair, start the one read-back of the frame
// (counters of every batch + the frame alpha counter: one copy of the whole block is cheaper than two small ones).
static int enqueue_frame_tail(pfcu_ctx *c) {
    int r = join_aux(c);  // batches that were prepared but not drawn (clip batches)
    if (r) return r;
    if (c->event_begin_recorded) CUDA_TRY(cudaEventRecord(c->ev_end, c->stream));
    CUDA_TRY(cudaMemcpyAsync(c->host_counters.p, c->counters.p, sizeof(BatchCounters) * (MAX_SLOTS + 1), cudaMemcpyDeviceToHost,
                             c->stream));
    return PFCU_OK;
}

int pfcu_submit_frame(pfcu_ctx *c) {
    NvtxScope nvtx("pfcu_submit_frame");
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    if (!c->frame_open) return fail(PFCU_ERR_STATE, "no frame is open");
    if (c->frame_pending) return fail(PFCU_ERR_STATE, "the frame has already been submitted");
    CUDA_TRY(cudaSetDevice(c->device));
    const uint64_t sig = c->frame_sig;
    bool served_by_graph = false;
    if (c->dry) {  // nothing has been enqueued yet: the retained graph serves the frame if it is the same frame
        c->dry = false;
        if (!c->cmds.empty()) {
            CUDA_TRY(cudaEventRecord(c->ev_begin, c->stream));
            c->event_begin_recorded = true;
            if (c->retained_exec && sig == c->retained_sig) {
                CUDA_TRY(cudaGraphLaunch(c->retained_exec, c->stream));
                c->launches = c->retained_launches;
                c->in_flight = true;
                served_by_graph = true;
            } else {
                drop_retained(c);
                c->frame_fork = nullptr;
        CUDA_TRY(cudaMemsetAsync(frame_alpha_counter(c), 0, sizeof(BatchCounters), c->stream));
                for (int i = 0; i < c->slots_used; i++) c->slots[i].prepared = false;
                for (const Cmd &cmd : c->cmds) {
                    const int r = cmd.kind == CMD_PREPARE ? enqueue_prepare(c, cmd.slot) : enqueue_draw(c, cmd);
                    if (r) {
                        c->frame_open = false;
                        return r;
                    }
                }
            }
        }
    }
    {
        const int r = enqueue_frame_tail(c);
        if (r) {
            c->frame_open = false;
            return r;
        }
    }
    c->pending_served_by_graph = served_by_graph;
    c->pending_sig = sig;
    c->frame_pending = true;
    return PFCU_OK;
}

int pfcu_wait_frame(pfcu_ctx *c, pfcu_frame_stats *stats) {
    NvtxScope nvtx("pfcu_wait_frame");
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    if (!c->frame_pending) return fail(PFCU_ERR_STATE, "no frame has been submitted");
    CUDA_TRY(cudaSetDevice(c->device));
    c->frame_pending = false;
    BatchCounters *hc = static_cast<BatchCounters *>(c->host_counters.p);
    const uint64_t sig = c->pending_sig;
    const bool served_by_graph = c->pending_served_by_graph;
    int attempts_used = 0;
    for (int attempt = 0;; attempt++) {
        attempts_used = attempt;
        if (attempt > 0) {
            const int r = enqueue_frame_tail(c);
            if (r) {
                c->frame_open = false;
                return r;
            }
        }
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        c->in_flight = false;
        c->stage_busy[0] = c->stage_busy[1] = false;
        uint32_t overflow = 0, dda_anomaly = 0;
        const uint32_t frame_alpha = *reinterpret_cast<uint32_t *>(hc + MAX_SLOTS);
        for (int i = 0; i < c->slots_used; i++) {
            BatchSlot &s = c->slots[i];
            dda_anomaly |= hc[i].overflow & OVF_DDA;  // not a capacity problem: reported, never retried
            overflow |= hc[i].overflow & ~(uint32_t)OVF_DDA;
            if (hc[i].n_lines > s.line_cap) s.line_cap = (uint32_t)round_up_pow2(hc[i].n_lines);
            if (hc[i].n_fills > s.fill_cap) s.fill_cap = (uint32_t)round_up_pow2(hc[i].n_fills);
            if (hc[i].n_staging > s.staging_cap) s.staging_cap = (uint32_t)round_up_pow2(hc[i].n_staging);
            if ((hc[i].overflow & (OVF_LINES | OVF_STAGING)) && s.fill_cap < s.staging_cap) s.fill_cap = s.staging_cap;
            // a line overflow hides the true fill count: make sure fills can grow with the lines
            if ((hc[i].overflow & OVF_LINES) && s.fill_cap < s.line_cap * 2) s.fill_cap = s.line_cap * 2;
        }
        if (frame_alpha > c->mask_cap) {
            overflow |= OVF_ALPHA;
            c->mask_cap = (uint32_t)round_up_pow2(frame_alpha);
            CUDA_TRY(c->masks.ensure((size_t)c->mask_cap * 256));
        }
        c->last_stats.overflow_flags = dda_anomaly;
        if (!overflow) break;
        if (attempt >= 3) {
            c->frame_open = false;
            return fail(PFCU_ERR_OVERFLOW, "ran out of space after %d attempts (flags 0x%x)", attempt + 1, overflow);
        }
        // grow and replay the recorded frame
        drop_retained(c);  // (its parameter blocks point into the buffers that are about to move)
        for (int i = 0; i < c->slots_used; i++) {  // ... and so do the lines an incremental frame would have kept
            c->slots[i].base.valid = false;
            c->slots[i].diced_all = true;
            c->slots[i].diced_segments = c->slots[i].desc.segment_count;
        }
        c->retries++;
        c->prof_used = 0;
        c->sync_used = 0;
        for (int i = 0; i < c->slots_used; i++) c->slots[i].prepared = false;
        CUDA_TRY(cudaEventRecord(c->ev_begin, c->stream));
        c->frame_fork = nullptr;
        CUDA_TRY(cudaMemsetAsync(frame_alpha_counter(c), 0, sizeof(BatchCounters), c->stream));
        for (const Cmd &cmd : c->cmds) {
            const int r = cmd.kind == CMD_PREPARE ? enqueue_prepare(c, cmd.slot) : enqueue_draw(c, cmd);
            if (r) {
                c->frame_open = false;
                return r;
            }
        }
        if (c->read_pending) {  // a read-back was enqueued behind the attempt that overflowed: read the replayed frame
            CUDA_TRY(cudaStreamSynchronize(c->copy_stream));
            const int r = pfcu_read_target_async(c, c->read_host, c->read_pitch);
            if (r) {
                c->frame_open = false;
                return r;
            }
        }
    }
    // Retain the frame: the second of two identical frames is captured, identical frames after it are one graph launch.
    if (c->auto_graph && !c->profiling && !served_by_graph && !c->retained_exec && attempts_used == 0 && sig &&
        sig == c->prev_sig && !c->cmds.empty()) {
        const int r = capture_frame(c, true, &c->retained_graph, &c->retained_exec, &c->retained_launches);
        if (r == PFCU_OK) c->retained_sig = sig;
        else drop_retained(c);
    }
    c->prev_sig = sig;
    pfcu_frame_stats st{};
    if (c->incremental) {
        for (int i = 0; i < c->slots_used; i++) {
            BatchSlot &s = c->slots[i];
            if (s.diced_all) {  // everything this slot holds was diced from the current scene: the new base
                s.base.valid = true;
                s.base.key = s.frame_key;
                s.base.scene_gen = c->scene_gen[s.desc.path_source ? 1 : 0];
                s.base.seq = s.frame_seq;
                s.base.n_lines = hc[i].n_lines;
                s.base.n_staging = hc[i].n_staging;
                s.base.n_long = hc[i].n_long;
            }
            st.diced_segments += s.diced_segments;
        }
    } else {
        for (int i = 0; i < c->slots_used; i++) st.diced_segments += c->slots[i].desc.segment_count;
    }
    st.uploaded_bytes = c->uploaded_bytes;
    c->uploaded_bytes = 0;
    st.batches = (uint32_t)c->slots_used;
    for (int i = 0; i < c->slots_used; i++) {
        st.segments += c->slots[i].desc.segment_count;
        st.lines += hc[i].n_lines;
        st.fills += hc[i].n_fills;
        st.dense_tiles += c->slots[i].desc.tile_count;
        st.listed_tiles += hc[i].n_list_entries;
        st.listed_after_cull += hc[i].n_listed;
        st.max_list_len = std::max(st.max_list_len, hc[i].max_list_len);
    }
    st.alpha_tiles = *reinterpret_cast<uint32_t *>(hc + MAX_SLOTS);
    st.fb_tiles = (uint32_t)(((c->target.width + TILE - 1) / TILE) * ((c->target.height + TILE - 1) / TILE));
    st.retries = c->retries;
    st.overflow_flags = c->last_stats.overflow_flags;
    st.kernel_launches = c->launches;
    if (c->event_begin_recorded) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end) == cudaSuccess) st.gpu_ms = ms;
    }
    for (int k = 0; k < PFCU_NUM_STAGES; k++) c->stage_ms[k] = 0.f;
    if (c->profiling) {
        for (size_t i = 1; i < c->prof_used; i++) {
            const int stage = c->prof_stage[i];
            if (stage < 0) continue;
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, c->prof_events[i - 1], c->prof_events[i]) == cudaSuccess) c->stage_ms[stage] += ms;
        }
    }
    c->last_stats = st;
    if (stats) *stats = st;
    c->frame_open = false;
    return PFCU_OK;
}

int pfcu_end_frame(pfcu_ctx *c, pfcu_frame_stats *stats) {
    const int r = pfcu_submit_frame(c);
    return r ? r : pfcu_wait_frame(c, stats);
}

int pfcu_set_option(pfcu_ctx *c, int option, int value) {
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    if (c->frame_open) return fail(PFCU_ERR_STATE, "options cannot change inside a frame");
    switch (option) {
        case PFCU_OPT_RETAIN_FRAME_GRAPH:
            c->auto_graph = value != 0;
            if (!c->auto_graph) drop_retained(c);
            return PFCU_OK;
        case PFCU_OPT_FILL_CULLED_TILES:
            c->fill_culled_tiles = value != 0;
            return PFCU_OK;
        case PFCU_OPT_FUSED_FILL:
            c->fused_fill = value != 0;
            return PFCU_OK;
        case PFCU_OPT_CONCURRENT_BATCHES:
            c->concurrent_batches = value != 0;
            drop_retained(c);
            return PFCU_OK;
        case PFCU_OPT_ORDER_TILE_GROUPS:
            c->order_groups = value < 0 ? 0 : value;
            return PFCU_OK;
        case PFCU_OPT_INCREMENTAL_DICE:
            c->incremental = value != 0;
            for (auto &sl : c->slots) sl.base.valid = false;
            return PFCU_OK;
        default:
            return fail(PFCU_ERR_INVALID, "unknown option %d", option);
    }
}

int pfcu_set_profiling(pfcu_ctx *c, int enabled) {
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    c->profiling = enabled != 0;
    return PFCU_OK;
}

int pfcu_set_timeline(pfcu_ctx *c, uint32_t capacity) {
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    if (c->frame_open) return fail(PFCU_ERR_STATE, "the timeline cannot change inside a frame");
    if (capacity && !timeline_compiled())
        return fail(PFCU_ERR_STATE, "this library was built without -DPFCU_TIMELINE: load lib/libpfcu_trace.so (make trace)");
    cudaSetDevice(c->device);
    int r = sync_if_in_flight(c);
    if (r) return r;
    c->timeline_cap = 0;
    if (capacity) {
        CUDA_TRY(c->timeline.ensure(64 + (size_t)capacity * sizeof(pfcu_timeline_record)));
        CUDA_TRY(cudaMemset(c->timeline.p, 0, 64));
        c->timeline_cap = capacity;
    }
    return PFCU_OK;  // (the kernels' parameter blocks change: a retained frame graph is rebuilt by the next frames)
}

int64_t pfcu_read_timeline(pfcu_ctx *c, pfcu_timeline_record *out, int64_t max_records) {
    if (!c || !c->timeline_cap) {
        fail(PFCU_ERR_STATE, "pfcu_set_timeline has not been called");
        return -1;
    }
    cudaSetDevice(c->device);
    uint32_t n = 0;
    if (cudaStreamSynchronize(c->stream) != cudaSuccess || cudaStreamSynchronize(c->aux_stream) != cudaSuccess ||
        cudaMemcpy(&n, c->timeline.p, 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
        fail(PFCU_ERR_CUDA, "pfcu_read_timeline: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    if (!out) return n;
    const int64_t take = std::min<int64_t>(std::min<int64_t>(n, c->timeline_cap), std::max<int64_t>(max_records, 0));
    if (take > 0 && cudaMemcpy(out, c->timeline.as<char>() + 64, (size_t)take * sizeof(pfcu_timeline_record), cudaMemcpyDeviceToHost) != cudaSuccess) {
        fail(PFCU_ERR_CUDA, "pfcu_read_timeline: %s", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    cudaMemset(c->timeline.p, 0, 64);
    return n;
}

int pfcu_get_stage_times(pfcu_ctx *c, float *ms, int n) {
    if (!c || !ms) return fail(PFCU_ERR_INVALID, "null argument");
    for (int k = 0; k < n; k++) ms[k] = k < PFCU_NUM_STAGES ? c->stage_ms[k] : 0.f;
    return PFCU_OK;
}

// ------------------------------------------------------------------------------------------------ CUDA graph

int pfcu_graph_capture(pfcu_ctx *c) {
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    if (c->frame_open) return fail(PFCU_ERR_STATE, "end the frame before capturing it");
    if (c->cmds.empty()) return fail(PFCU_ERR_STATE, "no frame has been recorded");
    CUDA_TRY(cudaSetDevice(c->device));
    int r = sync_if_in_flight(c);
    if (r) return r;
    if (c->graph_exec) cudaGraphExecDestroy(c->graph_exec);
    if (c->graph) cudaGraphDestroy(c->graph);
    c->graph_exec = nullptr;
    c->graph = nullptr;
    // (the batch metadata is resident, so no host memory is touched by this graph)
    return capture_frame(c, false, &c->graph, &c->graph_exec, &c->graph_launches_per_frame);
}

int pfcu_graph_launch(pfcu_ctx *c) {
    if (!c || !c->graph_exec) return fail(PFCU_ERR_STATE, "no captured frame");
    CUDA_TRY(cudaGraphLaunch(c->graph_exec, c->stream));
    c->in_flight = true;
    return PFCU_OK;
}

int pfcu_graph_finish(pfcu_ctx *c, pfcu_frame_stats *stats) {
    if (!c || !c->graph_exec) return fail(PFCU_ERR_STATE, "no captured frame");
    CUDA_TRY(cudaSetDevice(c->device));
    BatchCounters *hc = static_cast<BatchCounters *>(c->host_counters.p);
    const size_t used = sizeof(BatchCounters) * (size_t)std::max(c->slots_used, 1);
    CUDA_TRY(cudaMemcpyAsync(hc, c->counters.p, used, cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaMemcpyAsync(hc + MAX_SLOTS, frame_alpha_counter(c), sizeof(BatchCounters), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->in_flight = false;
    pfcu_frame_stats st = c->last_stats;
    st.lines = st.fills = st.listed_tiles = st.listed_after_cull = st.max_list_len = 0;
    uint32_t overflow = 0;
    for (int i = 0; i < c->slots_used; i++) {
        overflow |= hc[i].overflow;
        st.lines += hc[i].n_lines;
        st.fills += hc[i].n_fills;
        st.listed_tiles += hc[i].n_list_entries;
        st.listed_after_cull += hc[i].n_listed;
        st.max_list_len = std::max(st.max_list_len, hc[i].max_list_len);
    }
    st.alpha_tiles = *reinterpret_cast<uint32_t *>(hc + MAX_SLOTS);
    if (st.alpha_tiles > c->mask_cap) overflow |= OVF_ALPHA;
    st.overflow_flags = overflow;
    st.kernel_launches = c->graph_launches_per_frame;
    st.retries = 0;
    if (stats) *stats = st;
    if (overflow) return fail(PFCU_ERR_OVERFLOW, "captured frame overflowed (flags 0x%x): re-run it with pfcu_end_frame", overflow);
    return PFCU_OK;
}

int pfcu_read_target(pfcu_ctx *c, uint8_t *host) {
    if (!c || !host || !c->target.pixels) return fail(PFCU_ERR_INVALID, "no target");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy2D(host, (size_t)c->target.width * 4, c->target.pixels, c->target.pitch,
                          (size_t)c->target.width * 4, c->target.height, cudaMemcpyDeviceToHost));
    return PFCU_OK;
}

// CommandEncoder::read_texture + the fence wait of Queue::submit_and_wait (gpu/command_encoder.cpp:317-355,
// gpu/vk/queue.cpp:29-35), split in two so that nothing blocks: the copy is enqueued on the context's copy stream behind
// the frame that is in flight (submitted or not yet waited for) and runs on a copy engine while the SMs render the next
// frame of ANOTHER context; a new frame on THIS context waits for the copy before it overwrites the target.
int pfcu_read_target_async(pfcu_ctx *c, uint8_t *host, size_t host_pitch_bytes) {
    if (!c || !host || !c->target.pixels) return fail(PFCU_ERR_INVALID, "no target");
    if (c->frame_open && !c->frame_pending) return fail(PFCU_ERR_STATE, "submit or end the frame before reading it back");
    const size_t row = (size_t)c->target.width * 4;
    if (!host_pitch_bytes) host_pitch_bytes = row;
    if (host_pitch_bytes < row) return fail(PFCU_ERR_INVALID, "host pitch smaller than a row");
    CUDA_TRY(cudaSetDevice(c->device));
    NvtxScope nvtx("pfcu_read_target_async");
    CUDA_TRY(cudaEventRecord(c->ev_frame_done, c->stream));
    CUDA_TRY(cudaStreamWaitEvent(c->copy_stream, c->ev_frame_done, 0));
    CUDA_TRY(cudaMemcpy2DAsync(host, host_pitch_bytes, c->target.pixels, c->target.pitch, row, c->target.height,
                               cudaMemcpyDeviceToHost, c->copy_stream));
    CUDA_TRY(cudaEventRecord(c->ev_copy_done, c->copy_stream));
    c->read_pending = true;
    c->read_host = host;
    c->read_pitch = host_pitch_bytes;
    return PFCU_OK;
}

int pfcu_wait_read(pfcu_ctx *c) {
    if (!c) return fail(PFCU_ERR_INVALID, "ctx is null");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaEventSynchronize(c->ev_copy_done));  // (returns at once when no read was ever enqueued)
    c->read_pending = false;
    return PFCU_OK;
}

// Page-locked host memory for pfcu_read_target_async / the upload calls (a pageable destination makes the copy
// synchronous and three times slower). Plain C callers need no CUDA headers.
void *pfcu_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        fail(PFCU_ERR_OOM, "cudaMallocHost(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}

void pfcu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int pfcu_read_target_region(pfcu_ctx *c, int x, int y, int width, int height, uint8_t *host) {
    if (!c || !host || !c->target.pixels) return fail(PFCU_ERR_INVALID, "no target");
    // CommandEncoder::read_texture (gpu/command_encoder.cpp:317-325): an invalid or empty region is an error
    if (x < 0 || y < 0 || width <= 0 || height <= 0 || x + width > c->target.width || y + height > c->target.height)
        return fail(PFCU_ERR_INVALID, "region %d,%d %dx%d is outside the %dx%d target", x, y, width, height,
                    c->target.width, c->target.height);
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy2D(host, (size_t)width * 4, c->target.pixels + (size_t)y * c->target.pitch + (size_t)x * 4,
                          c->target.pitch, (size_t)width * 4, height, cudaMemcpyDeviceToHost));
    return PFCU_OK;
}

int pfcu_read_page(pfcu_ctx *c, uint32_t page, uint8_t *host) {
    if (!c || !host || page >= MAX_PAGES || !c->pages[page].px.p) return fail(PFCU_ERR_INVALID, "texture page not allocated");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    CUDA_TRY(cudaMemcpy(host, c->pages[page].px.p, (size_t)c->pages[page].w * c->pages[page].h * 4, cudaMemcpyDeviceToHost));
    return PFCU_OK;
}

void *pfcu_target_device_ptr(pfcu_ctx *c, size_t *pitch) {
    if (!c) return nullptr;
    if (pitch) *pitch = c->target.pitch;
    return c->target.pixels;
}

// ------------------------------------------------------------------------------------------------ parity taps

static int tap_slot(pfcu_ctx *c, uint32_t batch_id) {
    if (!c) {
        fail(PFCU_ERR_INVALID, "ctx is null");
        return -1;
    }
    for (int i = 0; i < c->slots_used; i++)
        if (c->slots[i].desc.batch_id == batch_id) return i;
    fail(PFCU_ERR_INVALID, "batch %u not in the last frame", batch_id);
    return -1;
}

#define TAP_TRY(expr)                                                           \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) {                                                \
            fail(PFCU_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e));       \
            return -1;                                                          \
        }                                                                       \
    } while (0)

int64_t pfcu_read_lines(pfcu_ctx *c, uint32_t batch_id, pfcu_line *out) {
    const int si = tap_slot(c, batch_id);
    if (si < 0) return -1;
    cudaSetDevice(c->device);
    TAP_TRY(cudaStreamSynchronize(c->stream));
    BatchSlot &s = c->slots[si];
    BatchCounters bc;
    TAP_TRY(cudaMemcpy(&bc, s.view.counters, sizeof(bc), cudaMemcpyDeviceToHost));
    const uint32_t n = std::min(bc.n_lines, s.line_cap);
    if (!out) return n;
    std::vector<float4> l(n);
    std::vector<uint2> p(n);
    if (n) {
        TAP_TRY(cudaMemcpy(l.data(), s.view.lines, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost));
        TAP_TRY(cudaMemcpy(p.data(), s.view.line_meta, (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost));
    }
    for (uint32_t i = 0; i < n; i++) out[i] = pfcu_line{l[i].x, l[i].y, l[i].z, l[i].w, p[i].x};
    return n;
}

int64_t pfcu_read_fills(pfcu_ctx *c, uint32_t batch_id, pfcu_fill *out) {
    const int si = tap_slot(c, batch_id);
    if (si < 0) return -1;
    cudaSetDevice(c->device);
    TAP_TRY(cudaStreamSynchronize(c->stream));
    BatchSlot &s = c->slots[si];
    BatchCounters bc;
    TAP_TRY(cudaMemcpy(&bc, s.view.counters, sizeof(bc), cudaMemcpyDeviceToHost));
    const uint32_t n = std::min(bc.n_fills, s.fill_cap);
    if (!out) return n;
    const uint32_t D = s.desc.tile_count;
    std::vector<uint32_t> word(D), cursor(D);
    std::vector<uint2> f(n);
    if (D) {
        TAP_TRY(cudaMemcpy(word.data(), s.view.tile_word, (size_t)D * 4, cudaMemcpyDeviceToHost));
        TAP_TRY(cudaMemcpy(cursor.data(), s.view.fill_cursor, (size_t)D * 4, cudaMemcpyDeviceToHost));
    }
    if (n) TAP_TRY(cudaMemcpy(f.data(), s.view.fills, (size_t)n * sizeof(uint2), cudaMemcpyDeviceToHost));
    size_t w = 0;
    for (uint32_t t = 0; t < D; t++) {
        const uint32_t cnt = word[t] & 0x00ffffffu, end = std::min(cursor[t], n);
        const uint32_t begin = end >= cnt ? end - cnt : 0;
        const size_t w0 = w;
        for (uint32_t k = begin; k < end && w < n; k++) {
            pfcu_fill q;
            q.tile_index = t;
            q.from_x = (uint16_t)(f[k].x & 0xffff);
            q.from_y = (uint16_t)(f[k].x >> 16);
            q.to_x = (uint16_t)(f[k].y & 0xffff);
            q.to_y = (uint16_t)(f[k].y >> 16);
            out[w++] = q;
        }
        std::sort(out + w0, out + w, [](const pfcu_fill &a, const pfcu_fill &b) {
            if (a.from_x != b.from_x) return a.from_x < b.from_x;
            if (a.from_y != b.from_y) return a.from_y < b.from_y;
            if (a.to_x != b.to_x) return a.to_x < b.to_x;
            return a.to_y < b.to_y;
        });
    }
    return (int64_t)w;
}

int64_t pfcu_read_tiles(pfcu_ctx *c, uint32_t batch_id, pfcu_tile *out) {
    const int si = tap_slot(c, batch_id);
    if (si < 0) return -1;
    cudaSetDevice(c->device);
    TAP_TRY(cudaStreamSynchronize(c->stream));
    BatchSlot &s = c->slots[si];
    const uint32_t D = s.desc.tile_count;
    if (!out) return D;
    BatchCounters bc;
    TAP_TRY(cudaMemcpy(&bc, s.view.counters, sizeof(bc), cudaMemcpyDeviceToHost));
    std::vector<uint32_t> word(D);
    std::vector<TileState> st(D);
    if (D) {
        TAP_TRY(cudaMemcpy(word.data(), s.view.tile_word, (size_t)D * 4, cudaMemcpyDeviceToHost));
        TAP_TRY(cudaMemcpy(st.data(), s.view.tile_state, (size_t)D * sizeof(TileState), cudaMemcpyDeviceToHost));
    }
    for (uint32_t t = 0; t < D; t++) {
        pfcu_tile q{};
        q.alpha_tile_id = st[t].alpha;
        q.clip_alpha_tile_id = (st[t].packed & (1u << 25)) ? st[t].clip_alpha : -1;
        q.fill_count = (int32_t)(word[t] & 0x00ffffffu);
        q.backdrop = (int8_t)(st[t].packed & 0xff);
        q.backdrop_delta = (int8_t)((st[t].packed >> 8) & 0xff);
        q.backdrop_d3d9 = (int8_t)((st[t].packed >> 16) & 0xff);
        q.listed = (uint8_t)((st[t].packed >> 24) & 1);
        out[t] = q;
    }
    return D;
}

int64_t pfcu_read_z(pfcu_ctx *c, uint32_t batch_id, int32_t *out) {
    const int si = tap_slot(c, batch_id);
    if (si < 0) return -1;
    cudaSetDevice(c->device);
    TAP_TRY(cudaStreamSynchronize(c->stream));
    BatchSlot &s = c->slots[si];
    const uint32_t T = (uint32_t)(s.view.fb_tw * s.view.fb_th);
    if (out && T) {
        std::vector<FbTile> fb(T);
        TAP_TRY(cudaMemcpy(fb.data(), s.view.fb, (size_t)T * sizeof(FbTile), cudaMemcpyDeviceToHost));
        for (uint32_t t = 0; t < T; t++) out[t] = fb[t].z;
    }
    return T;
}

int64_t pfcu_read_tile_lists(pfcu_ctx *c, uint32_t batch_id, uint32_t *offsets, uint32_t *tiles) {
    const int si = tap_slot(c, batch_id);
    if (si < 0) return -1;
    cudaSetDevice(c->device);
    TAP_TRY(cudaStreamSynchronize(c->stream));
    BatchSlot &s = c->slots[si];
    const uint32_t T = (uint32_t)(s.view.fb_tw * s.view.fb_th), D = s.desc.tile_count;
    std::vector<FbTile> fb(T);
    std::vector<TilePrim> prims(D);
    if (T && s.view.fb_sorted) {  // the headers are in CTA order (PFCU_OPT_ORDER_TILE_GROUPS): bring them back to tile order
        const uint32_t groups = (T + GROUP_TILES - 1) / GROUP_TILES;
        std::vector<FbTile> sorted((size_t)groups * GROUP_TILES);
        std::vector<uint32_t> slot_of(groups);
        TAP_TRY(cudaMemcpy(sorted.data(), s.view.fb_sorted, sorted.size() * sizeof(FbTile), cudaMemcpyDeviceToHost));
        TAP_TRY(cudaMemcpy(slot_of.data(), s.view.slot_of, (size_t)groups * 4, cudaMemcpyDeviceToHost));
        for (uint32_t t = 0; t < T; t++) fb[t] = sorted[(size_t)slot_of[t / GROUP_TILES] * GROUP_TILES + t % GROUP_TILES];
    } else if (T)
        TAP_TRY(cudaMemcpy(fb.data(), s.view.fb, (size_t)T * sizeof(FbTile), cudaMemcpyDeviceToHost));
    if (D) TAP_TRY(cudaMemcpy(prims.data(), s.view.prims, (size_t)D * sizeof(TilePrim), cudaMemcpyDeviceToHost));
    int64_t total = 0;
    std::vector<uint32_t> keys;
    for (uint32_t t = 0; t < T; t++) {
        if (offsets) offsets[t] = (uint32_t)total;
        // the scatter only places what the z-buffer does not cull: `cursor` entries of the `count` slots are in use
        const uint32_t begin = std::min(fb[t].begin, D), end = std::min(fb[t].begin + std::min(fb[t].cursor, fb[t].count), D);
        keys.clear();
        for (uint32_t k = begin; k < end; k++) keys.push_back(prims[k].key & 0x00ffffffu);  // (high byte: LayerFlags)
        std::sort(keys.begin(), keys.end());
        if (tiles) memcpy(tiles + total, keys.data(), keys.size() * 4);
        total += (int64_t)keys.size();
    }
    if (offsets) offsets[T] = (uint32_t)total;
    return total;
}

int pfcu_read_mask(pfcu_ctx *c, uint32_t alpha_tile_id, uint8_t out[256]) {
    if (!c || !out) return fail(PFCU_ERR_INVALID, "null argument");
    if (alpha_tile_id >= c->mask_cap) return fail(PFCU_ERR_INVALID, "alpha tile id out of range");
    CUDA_TRY(cudaSetDevice(c->device));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    uint8_t raw[256];
    CUDA_TRY(cudaMemcpy(raw, c->masks.as<uint8_t>() + (size_t)alpha_tile_id * 256, 256, cudaMemcpyDeviceToHost));
    // device layout is lane-major (pfcu_device.h): byte lane * 8 + q = pixel (column (lane & 3) * 4 + (q & 3),
    // row (lane >> 2) * 2 + (q >> 2))
    for (int lane = 0; lane < 32; lane++)
        for (int q = 0; q < 8; q++) out[((lane >> 2) * 2 + (q >> 2)) * 16 + (lane & 3) * 4 + (q & 3)] = raw[lane * 8 + q];
    return PFCU_OK;
}

}  // extern "C"
