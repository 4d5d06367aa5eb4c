// bound / scan / propagate / sort: the integer bookkeeping between bin and the rasterizing kernels.
//
//   init ("bound") : pathfinder/shaders/d3d11/bound.comp:37-73 initialises 16 B per dense tile with a binary
//                    search per tile; here a tile is one 32-bit word (fill count | backdrop delta << 24) that only
//                    needs zeroing, the per-path control word is looked up where it is used. The same kernel
//                    resets the column backdrops, z-buffer, list counters and scan descriptors that the reference
//                    uploads from the CPU every batch (d3d11/renderer.cpp:565,868-883).
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
#include "pfcu_device.h"

namespace pfcu {

static int g_sm_count = 0;
int sm_count() {
    if (!g_sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__host__ __device__ static inline uint32_t scan_tiles_for(uint32_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// ------------------------------------------------------------------------------------------------ init

__global__ void __launch_bounds__(256) k_init(BatchView b) {
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t fbt = (uint32_t)(b.fb_tw * b.fb_th);
    for (uint32_t i = tid; i < b.tile_count; i += stride) b.tile_word[i] = 0u;
    for (uint32_t i = tid; i < b.column_count; i += stride) b.col_backdrop[i] = __ldg(&b.backdrops[i].initial_backdrop);
    for (uint32_t i = tid; i < fbt; i += stride) *reinterpret_cast<uint4 *>(&b.fb[i]) = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t i = tid; i < scan_tiles_for(b.tile_count); i += stride) b.scan_desc[0][i] = 0ull;
    for (uint32_t i = tid; i < scan_tiles_for(fbt); i += stride) b.scan_desc[1][i] = 0ull;
    if (tid == 0) {
        BatchCounters c = {};
        c.first_alpha = *b.frame_alpha_counter;  // batches of a frame are stream-ordered
        *b.counters = c;
    }
}

cudaError_t launch_init(const BatchView &b, cudaStream_t s) {
    const uint32_t fbt = (uint32_t)(b.fb_tw * b.fb_th);
    uint32_t n = b.tile_count > fbt ? b.tile_count : fbt;
    if (b.column_count > n) n = b.column_count;
    int grid = (int)((n + 255) / 256);
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    k_init<<<grid, 256, 0, s>>>(b);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ scan

constexpr unsigned long long FLAG_AGGREGATE = 1ull << 62, FLAG_PREFIX = 2ull << 62, FLAG_MASK = 3ull << 62;

// WHICH 0: fills per dense tile (tile_word & 0xffffff -> fill_cursor), 1: list entries per framebuffer tile
// (fb[t].count -> fb[t].begin).
template <int WHICH>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan(BatchView b) {
    const uint32_t n = WHICH == 0 ? b.tile_count : (uint32_t)(b.fb_tw * b.fb_th);
    unsigned long long *desc = b.scan_desc[WHICH];
    __shared__ uint32_t s_tile, s_warp[SCAN_THREADS / 32], s_prefix;
    if (threadIdx.x == 0) s_tile = atomicAdd(&b.counters->scan_ticket[WHICH], 1u);  // forward progress: tiles start in order
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t base = tile * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        uint32_t x = 0;
        if (base + k < n) x = WHICH == 0 ? (b.tile_word[base + k] & 0x00ffffffu) : b.fb[base + k].count;
        v[k] = sum;  // exclusive within the thread
        sum += x;
    }
    // block exclusive scan of the per-thread sums
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_off = 0, block_total = 0;
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
        uint32_t prefix = 0;
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
                uint32_t contrib = (idx >= 0 && (first < 0 || (int)lane <= first)) ? (uint32_t)(d & 0xffffffffull) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
                prefix += contrib;
                if (first >= 0) break;
                look -= 32;
            }
            if (lane == 0) atomicExch(&desc[tile], FLAG_PREFIX | (unsigned long long)(prefix + block_total));
        }
        if (lane == 0) s_prefix = prefix;
    }
    __syncthreads();
    const uint32_t off = s_prefix + warp_off + (incl - sum);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (base + k < n) {
            if (WHICH == 0) b.fill_cursor[base + k] = off + v[k];
            else b.fb[base + k].begin = off + v[k];
        }
    }
    // the last tile owns the grand total
    if (tile == scan_tiles_for(n) - 1 && threadIdx.x == 0) {
        const uint32_t total = s_prefix + block_total;
        if (WHICH == 0) {
            b.counters->n_fills = total;
            if (total > b.fill_capacity) atomicOr(&b.counters->overflow, (uint32_t)OVF_FILLS);
        } else {
            b.counters->n_list_entries = total;
            if (total > b.prim_capacity) atomicOr(&b.counters->overflow, (uint32_t)OVF_LIST);
        }
    }
}

cudaError_t launch_scan_tiles(const BatchView &b, cudaStream_t s) {
    if (!b.tile_count) return cudaSuccess;
    k_scan<0><<<scan_tiles_for(b.tile_count), SCAN_THREADS, 0, s>>>(b);
    return cudaGetLastError();
}

cudaError_t launch_scan_fb(const BatchView &b, cudaStream_t s) {
    const uint32_t fbt = (uint32_t)(b.fb_tw * b.fb_th);
    if (!fbt) return cudaSuccess;
    k_scan<1><<<scan_tiles_for(fbt), SCAN_THREADS, 0, s>>>(b);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ fill scatter

__global__ void __launch_bounds__(256) k_fill_scatter(BatchView b) {
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
    k_fill_scatter<<<sm_count() * 8, 256, 0, s>>>(b);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ propagate

// One warp per tile column; lanes are 32 consecutive rows (propagate.comp:95-216, tiler.cpp:369-439).
__global__ void __launch_bounds__(128) k_propagate(BatchView b) {
    const uint32_t col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned lane = threadIdx.x & 31;
    if (col >= b.column_count) return;
    const uint32_t path = __ldg(&b.backdrops[col].path_index);
    const int tx = __ldg(&b.backdrops[col].tile_x_offset);
    const int4 rect = __ldg(reinterpret_cast<const int4 *>(&b.meta[path].tile_rect[0]));
    const uint32_t tile_offset = __ldg(&b.meta[path].tile_offset);
    const int w = rect.z - rect.x, h = rect.w - rect.y;
    if (w <= 0 || h <= 0 || tx >= w) return;
    const pfcu_tile_path_info info = b.tpi[path];
    const uint32_t ctrl_base = (uint32_t)info.color | ((uint32_t)info.ctrl << 16);
    const int gx = tx + rect.x;
    const uint32_t z_write_path = __ldg(&b.meta[path].z_write);
    const uint32_t clip_index = __ldg(&b.meta[path].clip_path_index);
    const bool has_clip = (int32_t)clip_index >= 0;
    int4 crect = make_int4(0, 0, 0, 0);
    uint32_t ctile_offset = 0;
    const bool clip_ok = has_clip && b.clip_meta && clip_index < b.clip_path_count;
    if (clip_ok) {
        crect = __ldg(reinterpret_cast<const int4 *>(&b.clip_meta[clip_index].tile_rect[0]));
        ctile_offset = __ldg(&b.clip_meta[clip_index].tile_offset);
    }
    const bool even_odd = (info.ctrl & 0x2) != 0;
    const uint32_t first_alpha = b.counters->first_alpha;
    int carry = b.col_backdrop[col];

    for (int ty0 = 0; ty0 < h; ty0 += 32) {
        const int ty = ty0 + (int)lane;
        const bool valid = ty < h;
        const uint32_t ti = tile_offset + (uint32_t)tx + (uint32_t)w * (uint32_t)(valid ? ty : 0);
        const uint32_t word = valid ? b.tile_word[ti] : 0u;
        const int delta = (int)(int8_t)(word >> 24);
        // exclusive prefix of the deltas down the column (tiler.cpp:394,437)
        int incl = delta;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= (unsigned)d) incl += t;
        }
        const int cur = carry + incl - delta;
        carry += __shfl_sync(0xffffffffu, incl, 31);

        const bool have_mask = (word & 0x00ffffffu) != 0;
        int backdrop = (int)(int8_t)cur;  // int8_t(backdrops[column]), tiler.cpp:394
        int backdrop9 = backdrop;
        bool need_new = valid && have_mask;
        int alpha = -1, clip_alpha = -1;
        const int gy = ty + rect.y;
        if (has_clip && valid) {
            const bool inside = clip_ok && gx >= crect.x && gx < crect.z && gy >= crect.y && gy < crect.w;
            if (inside) {
                const TileState ct = b.clip_tile_state[ctile_offset + (uint32_t)(gx - crect.x) +
                                                       (uint32_t)(crect.z - crect.x) * (uint32_t)(gy - crect.y)];
                if (ct.alpha >= 0) {
                    if (have_mask) {  // tiler.cpp:403-414 / propagate.comp:144-147
                        clip_alpha = ct.alpha;
                        backdrop9 = 0;
                    } else if (backdrop != 0) {  // tiler.cpp:415-420 / propagate.comp:149-154
                        alpha = ct.alpha;
                        need_new = false;
                        backdrop9 = (int)(int8_t)((ct.packed >> 16) & 0xffu);
                    } else {
                        need_new = false;
                    }
                } else if ((int8_t)(ct.packed & 0xffu) == 0) {  // blank clip tile: tiler.cpp:421-425
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
        // alpha tile allocation (propagate.comp:178-183): one atomic per warp
        const unsigned need_mask = __ballot_sync(0xffffffffu, need_new);
        if (need_mask) {
            uint32_t base = 0;
            const int leader = __ffs(need_mask) - 1;
            if ((int)lane == leader) base = atomicAdd(b.frame_alpha_counter, (uint32_t)__popc(need_mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (need_new) {
                const uint32_t id = base + (uint32_t)__popc(need_mask & ((1u << lane) - 1u));
                const uint32_t local = id - first_alpha;
                if (id < b.mask_capacity && local < b.alpha_capacity) {
                    AlphaTile at;
                    at.tile_index = ti;
                    at.clip_alpha = clip_alpha;
                    at.packed = ((uint32_t)backdrop & 0xffu) | ((info.ctrl & 0x1) ? 0x100u : 0u);
                    at.fill_count = word & 0x00ffffffu;
                    *reinterpret_cast<uint4 *>(&b.alpha_tiles[local]) = *reinterpret_cast<const uint4 *>(&at);
                    alpha = (int)id;
                } else {
                    atomicOr(&b.counters->overflow, (uint32_t)OVF_ALPHA);
                    need_new = false;
                }
            }
        }
        const bool in_fb = valid && gx >= 0 && gx < b.fb_tw && gy >= 0 && gy < b.fb_th;
        const uint32_t map = in_fb ? (uint32_t)gy * (uint32_t)b.fb_tw + (uint32_t)gx : 0u;
        const bool listed = (backdrop != 0 || alpha >= 0) && in_fb;
        if (valid) {
            TileState st;
            st.alpha = alpha;
            st.packed = ((uint32_t)backdrop & 0xffu) | (((uint32_t)delta & 0xffu) << 8) |
                        (((uint32_t)backdrop9 & 0xffu) << 16) | (listed ? 1u << 24 : 0u) | (need_new ? 1u << 25 : 0u) |
                        (((uint32_t)info.ctrl & 0x3u) << 26);
            b.tile_state[ti] = st;
        }
        // z-buffer: propagate.comp:190-206 (even-odd tiles with an even backdrop are invisible, not occluders)
        bool z_write = z_write_path != 0;
        if (backdrop != 0 && even_odd && (abs(backdrop) & 1) == 0) z_write = false;
        if (in_fb && z_write && backdrop != 0 && alpha < 0) atomicMax(&b.fb[map].z, (int)ti);
        // list membership (propagate.comp:209-212): rank inside the framebuffer tile + a compact record
        const unsigned listed_mask = __ballot_sync(0xffffffffu, listed);
        if (listed_mask) {
            uint32_t base = 0;
            const int leader = __ffs(listed_mask) - 1;
            if ((int)lane == leader) base = atomicAdd(&b.counters->n_listed, (uint32_t)__popc(listed_mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (listed) {
                const uint32_t at = base + (uint32_t)__popc(listed_mask & ((1u << lane) - 1u));
                const uint32_t rank = atomicAdd(&b.fb[map].count, 1u);
                if (at < b.prim_capacity) {
                    *reinterpret_cast<uint4 *>(&b.listed[at]) =
                        make_uint4(ti, (uint32_t)alpha, ctrl_base | (((uint32_t)backdrop & 0xffu) << 24), map);
                    b.listed_rank[at] = rank;
                }
            }
        }
    }
}

cudaError_t launch_propagate(const BatchView &b, cudaStream_t s) {
    if (!b.column_count) return cudaSuccess;
    const uint32_t warps_per_block = 4;
    k_propagate<<<(b.column_count + warps_per_block - 1) / warps_per_block, 128, 0, s>>>(b);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ list scatter

__global__ void __launch_bounds__(256) k_list_scatter(BatchView b) {
    const uint32_t n = min(b.counters->n_listed, b.prim_capacity);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4 r = *reinterpret_cast<const uint4 *>(&b.listed[i]);
        const uint32_t pos = b.fb[r.w].begin + b.listed_rank[i];
        if (pos < b.prim_capacity) *reinterpret_cast<uint4 *>(&b.prims[pos]) = make_uint4(r.x, r.y, r.z, 0u);
    }
}

cudaError_t launch_list_scatter(const BatchView &b, cudaStream_t s) {
    if (!b.column_count) return cudaSuccess;
    k_list_scatter<<<sm_count() * 8, 256, 0, s>>>(b);
    return cudaGetLastError();
}

}  // namespace pfcu
