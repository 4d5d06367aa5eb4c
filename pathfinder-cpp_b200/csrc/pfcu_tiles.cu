// bound / scan / propagate / sort: the integer bookkeeping between bin and the rasterizing kernels.
//
//   init ("bound") : pathfinder/shaders/d3d11/bound.comp:37-73 initialises 16 B per dense tile with a binary
//                    search per tile; here a tile is one 32-bit word (fill count | backdrop delta << 24) that only
//                    needs zeroing, the per-path control word is looked up where it is used. The same kernel
//                    resets the column backdrops, z-buffer, list counters and scan descriptors that the reference
//                    uploads from the CPU every batch (d3d11/renderer.cpp:565,868-883). It runs beside dice.
//   scan           : single-pass decoupled look-back exclusive scan (Merrill & Garland 2016). It replaces the
//                    reference's atomic bump allocation + CPU read-back + retry (renderer.cpp:559-577,832-845):
//                    fill offsets per dense tile, list offsets per framebuffer tile.
//   fill scatter   : one thread per staged fill moves it to its tile's contiguous range (CSR).
//   propagate      : pathfinder/shaders/d3d11/propagate.comp:95-216 (== Tiler::prepare_tiles,
//                    core/d3d9/tiler.cpp:369-439). The reference walks each tile column serially in one thread;
//                    here a WARP owns a column, lanes are 32 consecutive rows, and the backdrop is a warp-shuffle
//                    prefix sum carried across 32-row chunks. Each lane then resolves its own tile: clip cases,
//                    alpha-tile allocation (one atomic per warp), z-buffer, list membership.
//   list scatter   : replaces the per-framebuffer-tile linked list (propagate.comp:209-212) + the global-memory
//                    insertion sort (sort.comp:49-83) with contiguous (CSR) lists: propagate takes a rank per
//                    listed tile, the scan turns counts into offsets, one thread per listed tile writes its entry.
//                    Ordering and z-culling happen on chip in the composite kernel.
#include <cuda_fp16.h>

#include "pfcu_device.h"

namespace pfcu {

// Per DEVICE, not per process: a process may hold contexts on several GPUs (pfcu_create(device_ordinal)).
int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return dev < 0 ? 0 : (dev >= MAX_DEVICES ? MAX_DEVICES - 1 : dev);
}

int sm_count() {
    static int counts[MAX_DEVICES] = {};
    const int dev = current_device();
    if (!counts[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        counts[dev] = n > 0 ? n : 148;
    }
    return counts[dev];
}

bool timeline_compiled() {
#ifdef PFCU_TIMELINE
    return true;
#else
    return false;
#endif
}

#ifndef SCAN_THREADS_N
#define SCAN_THREADS_N 256
#endif
constexpr int SCAN_THREADS = SCAN_THREADS_N;  // (>= 256: the descriptor arrays are sized for 2048 items per tile)
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__host__ __device__ static inline uint32_t scan_tiles_for(uint32_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// ------------------------------------------------------------------------------------------------ init

__global__ void __launch_bounds__(256) k_init(InitArgs b) {
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_INIT);
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t fbt = (uint32_t)(b.fb_tw * b.fb_th);
    for (uint32_t i = tid; i < b.tile_count; i += stride) b.tile_word[i] = 0u;
    for (uint32_t i = tid; i < b.column_count; i += stride) b.col_backdrop[i] = __ldg(&b.backdrops[i].initial_backdrop);
    for (uint32_t i = tid; i < fbt; i += stride) *reinterpret_cast<uint4 *>(&b.fb[i]) = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t i = tid; i < scan_tiles_for(b.tile_count); i += stride) b.scan_desc[0][i] = 0ull;
    for (uint32_t i = tid; i < scan_tiles_for(fbt); i += stride) b.scan_desc[1][i] = 0ull;
    // (the batch counters are zeroed by a memset node ahead of dice, which runs beside this kernel)
}

cudaError_t launch_init(const BatchView &b, cudaStream_t s) {
    const uint32_t fbt = (uint32_t)(b.fb_tw * b.fb_th);
    uint32_t n = b.tile_count > fbt ? b.tile_count : fbt;
    if (b.column_count > n) n = b.column_count;
    int grid = (int)((n + 255) / 256);
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    return launch_pdl(k_init, (unsigned)grid, 256, 0, s, InitArgs(b));
}

// ------------------------------------------------------------------------------------------------ scan

constexpr unsigned long long FLAG_AGGREGATE = 1ull << 62, FLAG_PREFIX = 2ull << 62, FLAG_MASK = 3ull << 62;
constexpr unsigned long long VALUE_MASK = ~FLAG_MASK;

// WHICH 0: per dense tile, fills (tile_word & 0xffffff -> fill_cursor) and, in the same pass, tiles that have fills
//          (-> alpha_rank: the tile's mask slot is first_alpha + rank, so propagate needs no allocation atomics).
//          The scanned value packs both: fills in bits 0-31, tiles in bits 32-55.
// WHICH 1: list entries per framebuffer tile (fb[t].count -> fb[t].begin). With ordered tile groups (BatchView::fb_sorted)
//          the same pass counts the EXPENSIVE groups (bits 32-55 of the scanned value; a group's flag rides on its first
//          tile): expensive group g goes to position (expensive groups before g), in grid order from the front, a cheap
//          one to (groups - 1 - cheap groups before g), from the back -- a stable two-way partition that needs no total.
//          The headers are written where the tile kernel's CTAs will look for them (fb_sorted).
template <int WHICH>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(ScanArgs b) {
    const uint32_t n = WHICH == 0 ? b.tile_count : (uint32_t)(b.fb_tw * b.fb_th);
    unsigned long long *desc = b.scan_desc[WHICH];
    __shared__ uint32_t s_tile;
    __shared__ unsigned long long s_warp[SCAN_THREADS / 32], s_prefix;
    PFCU_KERNEL_BEGIN(b, WHICH ? PFCU_STAGE_SCAN_FB : PFCU_STAGE_SCAN_TILES);
    const bool ordered = WHICH == 1 && b.fb_sorted != nullptr;
    if (threadIdx.x == 0) s_tile = atomicAdd(&b.counters->scan_ticket[WHICH], 1u);  // forward progress: tiles start in order
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    unsigned long long v[SCAN_ITEMS];
    unsigned long long sum = 0;
    uint32_t raw[SCAN_ITEMS];
    int32_t zs[WHICH == 1 ? SCAN_ITEMS : 1];
    if (WHICH == 0 && base + SCAN_ITEMS <= n) {  // 2 x 16-byte loads (SCAN_ITEMS == 8, base is a multiple of 8)
        const uint4 r0 = *reinterpret_cast<const uint4 *>(&b.tile_word[base]);
        const uint4 r1 = *reinterpret_cast<const uint4 *>(&b.tile_word[base + 4]);
        raw[0] = r0.x; raw[1] = r0.y; raw[2] = r0.z; raw[3] = r0.w;
        raw[4] = r1.x; raw[5] = r1.y; raw[6] = r1.z; raw[7] = r1.w;
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            raw[k] = 0;
            if (base + k < n) raw[k] = WHICH == 0 ? b.tile_word[base + k] : b.fb[base + k].count;
            if (WHICH == 1) zs[k] = (ordered && base + k < n) ? b.fb[base + k].z : 0;
        }
    }
    // (a thread's SCAN_ITEMS tiles are in one group: its first or its second half)
    static_assert(GROUP_TILES == 2 * SCAN_ITEMS, "a thread scans half a group of tiles");
    bool expensive = false;
    if (WHICH == 1 && ordered) {
        // propagate counted the masked entries of a tile in the high bits of its count: the group's cost is their sum over
        // its two halves (this thread's and its neighbour's)
        uint32_t masked = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            masked += raw[k] >> FB_COUNT_BITS;
            raw[k] &= (1u << FB_COUNT_BITS) - 1u;
        }
        masked += __shfl_xor_sync(0xffffffffu, masked, 1);
        expensive = masked >= b.group_cost_min;
    }
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        unsigned long long x = raw[k];
        if (WHICH == 0) {
            x &= 0x00ffffffull;
            if (x) x |= 1ull << 32;
        } else if (k == 0 && expensive && base % GROUP_TILES == 0) {
            x |= 1ull << 32;
        }
        v[k] = sum;  // exclusive within the thread
        sum += x;
    }
    // block exclusive scan of the per-thread sums
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned long long warp_off = 0, block_total = 0;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; w++) {
        if (w < (int)warp) warp_off += s_warp[w];
        block_total += s_warp[w];
    }
    // decoupled look-back by warp 0
    if (warp == 0) {
        if (lane == 0) {
            const unsigned long long d = (tile == 0 ? FLAG_PREFIX : FLAG_AGGREGATE) | block_total;
            atomicExch(&desc[tile], d);
        }
        unsigned long long prefix = 0;
        if (tile > 0) {
            int look = (int)tile - 1;
            while (true) {
                const int idx = look - (int)lane;
                unsigned long long d = FLAG_PREFIX;  // lanes past the start behave like a zero prefix
                if (idx >= 0) {
                    do {
                        d = *((volatile unsigned long long *)&desc[idx]);
                    } while ((d & FLAG_MASK) == 0);
                }
                const unsigned is_prefix = __ballot_sync(0xffffffffu, (d & FLAG_MASK) == FLAG_PREFIX);
                // add everything up to and including the closest PREFIX descriptor (lane 0 is the closest tile)
                const int first = __ffs(is_prefix) - 1;  // -1: all 32 predecessors only published aggregates
                unsigned long long contrib = (idx >= 0 && (first < 0 || (int)lane <= first)) ? (d & VALUE_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                prefix += contrib;
                if (first >= 0) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(&desc[tile], FLAG_PREFIX | (prefix + block_total));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    const unsigned long long off = s_prefix + warp_off + (incl - sum);
    uint32_t sorted_base = 0;
    if (WHICH == 1 && ordered && base < n) {
        const uint32_t g = base / GROUP_TILES, n_groups = (n + GROUP_TILES - 1) / GROUP_TILES;
        // expensive groups before g (the second half's prefix already holds the group's own flag)
        const uint32_t before = (uint32_t)(off >> 32) - ((base % GROUP_TILES) && expensive ? 1u : 0u);
        const uint32_t slot = expensive ? before : n_groups - 1 - (g - before);
        sorted_base = slot * GROUP_TILES + base % GROUP_TILES;
        if (base % GROUP_TILES == 0) {
            b.slot_of[g] = slot;
            b.group_of[slot] = g;
        }
    }
    if (WHICH == 0 && base + SCAN_ITEMS <= n) {
        uint32_t f[SCAN_ITEMS], a[SCAN_ITEMS];
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            const unsigned long long o = off + v[k];
            f[k] = (uint32_t)o;
            a[k] = (uint32_t)(o >> 32);
        }
        *reinterpret_cast<uint4 *>(&b.fill_begin[base]) = make_uint4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<uint4 *>(&b.fill_begin[base + 4]) = make_uint4(f[4], f[5], f[6], f[7]);
        *reinterpret_cast<uint4 *>(&b.fill_cursor[base]) = make_uint4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<uint4 *>(&b.fill_cursor[base + 4]) = make_uint4(f[4], f[5], f[6], f[7]);
        *reinterpret_cast<uint4 *>(&b.alpha_rank[base]) = make_uint4(a[0], a[1], a[2], a[3]);
        *reinterpret_cast<uint4 *>(&b.alpha_rank[base + 4]) = make_uint4(a[4], a[5], a[6], a[7]);
    } else {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            if (base + k < n) {
                const unsigned long long o = off + v[k];
                if (WHICH == 0) {
                    b.fill_begin[base + k] = (uint32_t)o;
                    b.fill_cursor[base + k] = (uint32_t)o;
                    b.alpha_rank[base + k] = (uint32_t)(o >> 32);
                } else if (ordered) {
                    // the whole header, where the CTA that renders this tile's group will look for it (cursor = 0)
                    *reinterpret_cast<uint4 *>(&b.fb_sorted[sorted_base + k]) = make_uint4((uint32_t)o, raw[k], (uint32_t)zs[k], 0u);
                } else {
                    b.fb[base + k].begin = (uint32_t)o;
                }
            }
        }
    }
    // the last tile owns the grand total
    if (tile == scan_tiles_for(n) - 1 && threadIdx.x == 0) {
        const unsigned long long total = s_prefix + block_total;
        if (WHICH == 0) {
            const uint32_t fills = (uint32_t)total, alphas = (uint32_t)(total >> 32);
            b.counters->n_fills = fills;
            b.counters->n_alpha = alphas;
            // the frame-global mask slot counter (the batches of a frame may prepare side by side: PFCU_OPT_CONCURRENT_BATCHES)
            const uint32_t first = atomicAdd(b.frame_alpha_counter, alphas);
            b.counters->first_alpha = first;  // read by propagate and fill, which run after this kernel
            uint32_t ovf = 0;
            if (fills > b.fill_capacity) ovf |= OVF_FILLS;
            if (first + alphas > b.mask_capacity || alphas > b.alpha_capacity) ovf |= OVF_ALPHA;
            if (ovf) atomicOr(&b.counters->overflow, ovf);
        } else {
            b.counters->n_list_entries = (uint32_t)total;
            if ((uint32_t)total > b.prim_capacity) atomicOr(&b.counters->overflow, (uint32_t)OVF_LIST);
        }
    }
}

cudaError_t launch_scan_tiles(const BatchView &b, cudaStream_t s) {
    if (!b.tile_count) return cudaSuccess;
    return launch_pdl(k_scan<0>, scan_tiles_for(b.tile_count), SCAN_THREADS, 0, s, ScanArgs(b));
}

cudaError_t launch_scan_fb(const BatchView &b, cudaStream_t s) {
    const uint32_t fbt = (uint32_t)(b.fb_tw * b.fb_th);
    if (!fbt) return cudaSuccess;
    return launch_pdl(k_scan<1>, scan_tiles_for(fbt), SCAN_THREADS, 0, s, ScanArgs(b));
}

// ------------------------------------------------------------------------------------------------ fill scatter

__global__ void __launch_bounds__(256) k_fill_scatter(FillScatterArgs b) {
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_FILL_SCATTER);
    const uint32_t n = min(b.counters->n_staging, b.staging_capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 f = *reinterpret_cast<const uint4 *>(&b.staging[i]);
        if (f.x >= b.tile_count) continue;  // unused slot
        const uint32_t pos = atomicAdd(&b.fill_cursor[f.x], 1u);
        if (pos < b.fill_capacity) b.fills[pos] = make_uint2(f.y, f.z);
    }
}

cudaError_t launch_fill_scatter(const BatchView &b, cudaStream_t s) {
    if (!b.segment_count || !b.tile_count) return cudaSuccess;
    return launch_pdl(k_fill_scatter, sm_count() * 8, 256, 0, s, FillScatterArgs(b));
}

// ------------------------------------------------------------------------------------------------ propagate

// What propagate needs to know about a tile column (everything but the running backdrop).
struct ColumnInfo {
    uint32_t path, tile_offset, z_write_path, clip_index, ctile_offset;
    int tx, w, h, gx, rect_y;
    int4 crect;
    uint32_t ctrl;  // TilePathInfo::ctrl
    bool has_clip, clip_ok;
};

template <class B>
__device__ __forceinline__ bool load_column(const B &b, uint32_t col, ColumnInfo &ci) {
    ci.path = __ldg(&b.backdrops[col].path_index);
    ci.tx = __ldg(&b.backdrops[col].tile_x_offset);
    const int4 rect = __ldg(reinterpret_cast<const int4 *>(&b.meta[ci.path].tile_rect[0]));
    ci.tile_offset = __ldg(&b.meta[ci.path].tile_offset);
    ci.w = rect.z - rect.x;
    ci.h = rect.w - rect.y;
    ci.gx = ci.tx + rect.x;
    ci.rect_y = rect.y;
    if (ci.w <= 0 || ci.h <= 0 || ci.tx >= ci.w) return false;
    ci.ctrl = b.tpi[ci.path].ctrl;
    ci.z_write_path = __ldg(&b.meta[ci.path].z_write);
    ci.clip_index = __ldg(&b.meta[ci.path].clip_path_index);
    ci.has_clip = (int32_t)ci.clip_index >= 0;
    ci.crect = make_int4(0, 0, 0, 0);
    ci.ctile_offset = 0;
    ci.clip_ok = ci.has_clip && b.clip_meta && ci.clip_index < b.clip_path_count;
    if (ci.clip_ok) {
        ci.crect = __ldg(reinterpret_cast<const int4 *>(&b.clip_meta[ci.clip_index].tile_rect[0]));
        ci.ctile_offset = __ldg(&b.clip_meta[ci.clip_index].tile_offset);
    }
    return true;
}

// One tile of a column (propagate.comp:118-213, tiler.cpp:391-437): `cur` is the column's backdrop BEFORE this tile's
// delta. No atomic returns a value: mask slots come from the scan, list positions are taken later by the list scatter.
template <class B>
__device__ __forceinline__ void resolve_tile(const B &b, const ColumnInfo &ci, uint32_t first_alpha, int ty, int cur,
                                             uint32_t word) {
    const uint32_t ti = ci.tile_offset + (uint32_t)ci.tx + (uint32_t)ci.w * (uint32_t)ty;
    const int delta = (int)(int8_t)(word >> 24);
    const bool even_odd = (ci.ctrl & 0x2) != 0;
    const uint32_t fill_count = word & 0x00ffffffu;
    const bool have_mask = fill_count != 0;
    int backdrop = (int)(int8_t)cur;  // int8_t(backdrops[column]), tiler.cpp:394
    int backdrop9 = backdrop;
    bool need_new = have_mask;
    int alpha = -1, clip_alpha = -1;
    const int gx = ci.gx, gy = ty + ci.rect_y;
    if (ci.has_clip) {
        const int4 crect = ci.crect;
        const bool inside = ci.clip_ok && gx >= crect.x && gx < crect.z && gy >= crect.y && gy < crect.w;
        if (inside) {
            const uint4 ct = *reinterpret_cast<const uint4 *>(
                &b.clip_tile_state[ci.ctile_offset + (uint32_t)(gx - crect.x) +
                                   (uint32_t)(crect.z - crect.x) * (uint32_t)(gy - crect.y)]);
            if ((int)ct.x >= 0) {
                if (have_mask) {  // tiler.cpp:403-414 / propagate.comp:144-147
                    clip_alpha = (int)ct.x;
                    backdrop9 = 0;
                } else if (backdrop != 0) {  // tiler.cpp:415-420 / propagate.comp:149-154
                    alpha = (int)ct.x;
                    need_new = false;
                    backdrop9 = (int)(int8_t)((ct.y >> 16) & 0xffu);
                } else {
                    need_new = false;
                }
            } else if ((int8_t)(ct.y & 0xffu) == 0) {  // blank clip tile: tiler.cpp:421-425
                backdrop = 0;
                backdrop9 = 0;
                need_new = false;
            }
        } else {  // outside the clip rect: tiler.cpp:426-430
            backdrop = 0;
            backdrop9 = 0;
            need_new = false;
        }
    }
    // alpha tile allocation (propagate.comp:178-183): the slot was fixed by the scan over tiles with fills
    uint32_t own_slot = 0xffffffffu;
    if (have_mask) {
        const uint32_t local = b.alpha_rank[ti];
        const uint32_t id = first_alpha + local;
        if (id < b.mask_capacity && local < b.alpha_capacity) {
            // a tile whose fills the clip made invisible keeps its slot but is marked so that fill skips it.
            // Record: tile | winding << 31, clip mask slot, first fill, backdrop | fill count << 8
            *reinterpret_cast<uint4 *>(&b.alpha_tiles[local]) =
                make_uint4((need_new ? ti : 0x7fffffffu) | ((ci.ctrl & 0x1) ? 0x80000000u : 0u), (uint32_t)clip_alpha,
                           b.fill_begin[ti], ((uint32_t)backdrop & 0xffu) | (fill_count << 8));
            own_slot = local;
            if (need_new) alpha = (int)id;
        } else {
            need_new = false;  // the scan flagged OVF_ALPHA: the frame is replayed with more slots
        }
    }
    const int fx = gx - b.fb_tx0, fy = gy - b.fb_ty0;  // framebuffer tile
    const bool in_fb = fx >= 0 && fx < b.fb_tw && fy >= 0 && fy < b.fb_th;
    const uint32_t map = in_fb ? (uint32_t)fy * (uint32_t)b.fb_tw + (uint32_t)fx : 0u;
    if (own_slot != 0xffffffffu) b.alpha_map[own_slot] = in_fb ? map : 0xffffffffu;  // for fill's z-cull
    const bool listed = (backdrop != 0 || alpha >= 0) && in_fb;
    const uint32_t packed = ((uint32_t)backdrop & 0xffu) | (((uint32_t)delta & 0xffu) << 8) |
                            (((uint32_t)backdrop9 & 0xffu) << 16) | (listed ? 1u << 24 : 0u) |
                            (need_new ? 1u << 25 : 0u) | ((ci.ctrl & 0x3u) << 26);
    *reinterpret_cast<uint4 *>(&b.tile_state[ti]) = make_uint4((uint32_t)alpha, packed, ci.path, (uint32_t)clip_alpha);
    // z-buffer: propagate.comp:190-206 (even-odd tiles with an even backdrop are invisible, not occluders)
    bool z_write = ci.z_write_path != 0;
    if (backdrop != 0 && even_odd && (abs(backdrop) & 1) == 0) z_write = false;
    if (in_fb && z_write && backdrop != 0 && alpha < 0) atomicMax(&b.fb[map].z, (int)ti);
    // list membership (propagate.comp:209-212): count now, place after the scan (fire-and-forget reduction)
    // (ordered tile groups: a masked entry also counts in the high bits -- the scan over framebuffer tiles adds them up per
    // group of tiles and splits the two numbers again)
    if (listed) atomicAdd(&b.fb[map].count, 1u + ((b.fb_sorted && alpha >= 0) ? 1u << FB_COUNT_BITS : 0u));
}

#ifndef PROPAGATE_SHORT_N
#define PROPAGATE_SHORT_N 4
#endif
constexpr int PROPAGATE_SHORT = PROPAGATE_SHORT_N;  // columns of up to this many tiles are walked by one thread

// propagate.comp:95-216 (== Tiler::prepare_tiles, tiler.cpp:369-439). One warp per tile column, in groups of 32
// consecutive columns:
//   * a TALL column is walked by its own warp: lanes are 32 consecutive rows and the backdrop is a warp-shuffle prefix
//     sum carried across 32-row chunks (the reference walks up to 256 tiles serially in one thread);
//   * the SHORT columns of a group (glyph-sized paths: the bulk of a text-density scene) are all walked by the group's
//     first warp, ONE LANE per column, serially, as the reference does -- neighbouring lanes are neighbouring columns of
//     the same path, so every row is one coalesced access and no lane idles below a 3-tile column. The other warps of
//     the group find their column short and leave.
__global__ void __launch_bounds__(128) k_propagate(PropagateArgs b) {
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_PROPAGATE);
    // this warp's column (dealing consecutive columns to consecutive CTAs instead was measured: no change, 16.4 us)
    const uint32_t wcol = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (wcol >= b.column_count) return;
    const uint32_t first_alpha = b.counters->first_alpha;
    {
        ColumnInfo cw = {};
        if (load_column(b, wcol, cw) && cw.h > PROPAGATE_SHORT) {  // (warp-uniform: every lane loaded the same column)
            int wcarry = b.col_backdrop[wcol];
            for (int ty0 = 0; ty0 < cw.h; ty0 += 32) {
                const int ty = ty0 + (int)lane;
                const bool in = ty < cw.h;
                const uint32_t word = in ? b.tile_word[cw.tile_offset + (uint32_t)cw.tx + (uint32_t)cw.w * (uint32_t)ty] : 0u;
                const int delta = (int)(int8_t)(word >> 24);
                // exclusive prefix of the deltas down the column (tiler.cpp:394,437)
                int incl = delta;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, d);
                    if (lane >= (unsigned)d) incl += t;
                }
                const int cur = wcarry + incl - delta;
                wcarry += __shfl_sync(0xffffffffu, incl, 31);
                if (in) resolve_tile(b, cw, first_alpha, ty, cur, word);
            }
        }
    }
    if (wcol & 31u) return;
    // the group's short columns, one per lane
    const uint32_t col = wcol + lane;
    ColumnInfo ci = {};
    if (col >= b.column_count || !load_column(b, col, ci) || ci.h > PROPAGATE_SHORT) return;
    int carry = b.col_backdrop[col];
    uint32_t words[PROPAGATE_SHORT];  // all loads first: the rows only depend on each other through the running backdrop
#pragma unroll
    for (int ty = 0; ty < PROPAGATE_SHORT; ty++)
        words[ty] = ty < ci.h ? b.tile_word[ci.tile_offset + (uint32_t)ci.tx + (uint32_t)ci.w * (uint32_t)ty] : 0u;
#pragma unroll
    for (int ty = 0; ty < PROPAGATE_SHORT; ty++) {
        if (ty < ci.h) resolve_tile(b, ci, first_alpha, ty, carry, words[ty]);
        carry += (int)(int8_t)(words[ty] >> 24);
    }
}

cudaError_t launch_propagate(const BatchView &b, cudaStream_t s) {
    if (!b.column_count) return cudaSuccess;
    const uint32_t warps_per_block = 4;
    return launch_pdl(k_propagate, (b.column_count + warps_per_block - 1) / warps_per_block, 128, 0, s, PropagateArgs(b));
}

// ------------------------------------------------------------------------------------------------ list scatter

// One thread per dense tile (coalesced 16-byte state loads): a listed tile that the z-buffer does not cull
// (sort.comp:62: tiles below the top-most occluder of their framebuffer tile) takes the next free position of its
// framebuffer tile's range and writes what the composite kernel needs about it as one 16-byte record. Culled tiles
// leave their slot unused: fb[].cursor ends up as the list length, fb[].count stays the slot count.
__global__ void __launch_bounds__(256) k_list_scatter(ListArgs b) {
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_LIST_SCATTER);
    const unsigned lane = threadIdx.x & 31;
    uint32_t placed = 0, longest = 0;
    for (uint32_t ti0 = blockIdx.x * blockDim.x; ti0 < b.tile_count; ti0 += gridDim.x * blockDim.x) {
        const uint32_t ti = ti0 + threadIdx.x;
        if (ti >= b.tile_count) continue;
        const uint4 st = *reinterpret_cast<const uint4 *>(&b.tile_state[ti]);
        if (!(st.y & (1u << 24))) continue;
        const uint32_t path = st.z;
        const int4 rect = __ldg(reinterpret_cast<const int4 *>(&b.meta[path].tile_rect[0]));
        const uint32_t local = ti - __ldg(&b.meta[path].tile_offset);
        const uint32_t w = (uint32_t)(rect.z - rect.x);
        const int fx = rect.x + (int)(local % w) - b.fb_tx0, fy = rect.y + (int)(local / w) - b.fb_ty0;
        const uint32_t map = (uint32_t)fy * (uint32_t)b.fb_tw + (uint32_t)fx;
        FbTile *const fbh = fb_header(b, map);
        const uint4 hdr = *reinterpret_cast<const uint4 *>(fbh);  // begin and z are final; cursor is moving
        if ((int)ti < (int)hdr.z) continue;
        const uint32_t pi = __ldg(reinterpret_cast<const uint32_t *>(&b.tpi[path]) + 3);  // color, ctrl, backdrop
        const uint32_t at = atomicAdd(&fbh->cursor, 1u);
        const uint32_t pos = hdr.x + at;
        placed++;
        longest = max(longest, at + 1u);
        if (pos >= b.prim_capacity) continue;
        uint4 rec;
        if (b.solid_prims) {
            // the layer, resolved: flags + the paint's base colour (the texels are halfs: nothing is lost)
            const uint32_t paint = pi & 0xffffu;
            float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
            if (paint < b.n_paints) c = __ldg(&b.paints[paint].base);
            const uint32_t fl = layer_flags((int)st.x, (pi >> 16) & 0xffu, (int)(int8_t)(st.y & 0xffu), b.mask_capacity);
            const __half2 rg = __floats2half2_rn(c.x, c.y), ba = __floats2half2_rn(c.z, c.w);
            rec = make_uint4(ti | (fl << 24), st.x, *reinterpret_cast<const uint32_t *>(&rg), *reinterpret_cast<const uint32_t *>(&ba));
        } else {
            rec = make_uint4(ti, st.x, (pi & 0x00ffffffu) | ((st.y & 0xffu) << 24), 0u);
        }
        *reinterpret_cast<uint4 *>(&b.prims[pos]) = rec;
    }
    // frame statistics (pfcu_frame_stats::listed_after_cull / max_list_len): one pair of atomics per CTA
    __shared__ uint32_t s_placed, s_longest;
    if (threadIdx.x == 0) s_placed = s_longest = 0;
    __syncthreads();
    placed = __reduce_add_sync(0xffffffffu, placed);
    longest = __reduce_max_sync(0xffffffffu, longest);
    if (lane == 0 && placed) {
        atomicAdd(&s_placed, placed);
        atomicMax(&s_longest, longest);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_placed) {
        atomicAdd(&b.counters->n_listed, s_placed);
        atomicMax(&b.counters->max_list_len, s_longest);
    }
}

cudaError_t launch_list_scatter(const BatchView &b, cudaStream_t s) {
    if (!b.column_count || !b.tile_count) return cudaSuccess;
    int grid = (int)((b.tile_count + 255) / 256);
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    return launch_pdl(k_list_scatter, (unsigned)grid, 256, 0, s, ListArgs(b));
}

}  // namespace pfcu
