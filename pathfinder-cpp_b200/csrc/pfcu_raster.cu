// fill + tile ("composite"): the two rasterizing stages, the HBM-bound end of the path.
//
// Pixel ownership is the same in both kernels so that they can be fused: a WARP owns one 16 x 16 tile, lane
// (c = lane & 15, h = lane >> 4) owns pixel column c, rows h*4 .. h*4+3 and 8+h*4 .. 8+h*4+3, i.e. two of the area
// LUT's 4-row groups, one in the top half of the tile and one in the bottom half (a short fill touches one half, so
// the other half's LUT fetches are skipped by the whole warp).
//
//   fill : pathfinder/shaders/d3d11/fill.comp:51-154. The tile's fills are contiguous (CSR from the scatter), read
//          once with coalesced 8-byte loads and broadcast by shuffle. The 256 x 256 area LUT sits behind the
//          texture unit (texel fetch + unorm conversion; the bilinear weights are applied in fp32); lanes whose
//          column a fill does not overlap (dX == 0, contribution exactly 0) skip it, and 4-row groups outside the
//          LUT's transition band take the saturated value without fetching. Coverage stays in registers; the mask
//          is stored once, lane-major, one 8-byte store per lane.
//   tile : pathfinder/shaders/d3d11/tile.comp:737-850 with the shading functions of tile.comp:126-134 (combine),
//          :319-347 (radial gradient), :354-392 (blur), :459-582 (composite), :586-607 (mask), :694-726 (paint
//          metadata, decoded once per upload into a float table). One 16-byte load gives the tile's list range and
//          z; the list (<= 32 entries: registers + shuffles, longer: selection from global memory) is ordered by
//          paint order and z-culled on chip (sort.comp:49-83). While every layer so far covered the whole tile with
//          one colour, the warp blends ONE pixel instead of 256 and stores the tile with 16-byte stores; the first
//          masked / textured layer expands it to per-pixel registers.
//   fused: in the fused instantiation the composite kernel computes the coverage of a draw tile from its fills
//          right where it is blended -- the mask never leaves the SM (B_fused). Clip masks
//          (written by the clip batch's fill kernel) are still read from memory and min()-ed in.
#include <cuda_fp16.h>

#include "pfcu_device.h"

namespace pfcu {

__device__ __forceinline__ float4 ld_rgba8(const uint8_t *px, int w, int x, int y) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(px) + (size_t)y * w + x);
    const float k = 1.0f / 255.0f;
    return make_float4((float)(v & 0xffu) * k, (float)((v >> 8) & 0xffu) * k, (float)((v >> 16) & 0xffu) * k,
                       (float)(v >> 24) * k);
}

__device__ __forceinline__ int wrap_or_clamp(int i, int n, bool repeat) {
    if (repeat) {
        i %= n;
        return i < 0 ? i + n : i;
    }
    return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// texture(sampler2D, uv) on an RGBA8 image: GL/Vulkan bilinear footprint, clamp-to-edge or repeat, or nearest.
__device__ __forceinline__ float4 sample_rgba8(const uint8_t *px, int w, int h, float u, float v, bool repeat_u,
                                               bool repeat_v, bool nearest) {
    float x = u * (float)w, y = v * (float)h;
    if (nearest) {
        return ld_rgba8(px, w, wrap_or_clamp((int)floorf(x), w, repeat_u), wrap_or_clamp((int)floorf(y), h, repeat_v));
    }
    x -= 0.5f;
    y -= 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float ax = x - fx0, ay = y - fy0;
    const int x0 = wrap_or_clamp((int)fx0, w, repeat_u), x1 = wrap_or_clamp((int)fx0 + 1, w, repeat_u);
    const int y0 = wrap_or_clamp((int)fy0, h, repeat_v), y1 = wrap_or_clamp((int)fy0 + 1, h, repeat_v);
    const float4 a = ld_rgba8(px, w, x0, y0), b = ld_rgba8(px, w, x1, y0);
    const float4 c = ld_rgba8(px, w, x0, y1), d = ld_rgba8(px, w, x1, y1);
    float4 r;
    {
        const float t0 = a.x + (b.x - a.x) * ax, t1 = c.x + (d.x - c.x) * ax;
        r.x = t0 + (t1 - t0) * ay;
    }
    {
        const float t0 = a.y + (b.y - a.y) * ax, t1 = c.y + (d.y - c.y) * ax;
        r.y = t0 + (t1 - t0) * ay;
    }
    {
        const float t0 = a.z + (b.z - a.z) * ax, t1 = c.z + (d.z - c.z) * ax;
        r.z = t0 + (t1 - t0) * ay;
    }
    {
        const float t0 = a.w + (b.w - a.w) * ax, t1 = c.w + (d.w - c.w) * ax;
        r.w = t0 + (t1 - t0) * ay;
    }
    return r;
}

__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float glsl_mod(float x, float y) { return x - y * floorf(x / y); }

// ------------------------------------------------------------------------------------------------ fill

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, unsigned lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += t;
    }
    return v;
}

// One 4-row group of one pixel column (computeCoverage's LUT fetch, fill.comp:70, times dX, into cov[0..3]).
// texture(uAreaLUT, uv) is spelled out in fp32: the texture unit fetches the four texels (point sampling,
// clamp-to-edge, unorm8 -> float) and the bilinear weights k** (already multiplied by dX) are applied here. The unit's
// own bilinear mode keeps only 8 fraction bits of the weights, which moves about 1 % of the mask bytes by one step --
// too coarse for a 1/255 bound on pixels under several translucent layers.
// Outside the LUT's transition band all four channels are exactly 1 (rows below the line) or exactly 0 (above it);
// `band` says the uploaded LUT has been checked for that (pfcu_set_area_lut), so when the whole warp is outside the
// band the group costs two compares instead of four fetches.
__device__ __forceinline__ void add_group(cudaTextureObject_t lut, bool band, float x, float fx0, float fy0, float half,
                                          float k00, float k10, float k01, float k11, float dX, float *cov) {
    int kind = 0;  // 0: sample, 1: all rows fully covered, 2: no row covered
    if (band) {
        if (x + 2.0f < 120.0f - half) kind = 1;
        else if (x - 1.0f > 184.0f + half) kind = 2;
    }
    if (__any_sync(__activemask(), kind == 0)) {
        const float4 t00 = tex2D<float4>(lut, fx0 + 0.5f, fy0 + 0.5f), t10 = tex2D<float4>(lut, fx0 + 1.5f, fy0 + 0.5f);
        const float4 t01 = tex2D<float4>(lut, fx0 + 0.5f, fy0 + 1.5f), t11 = tex2D<float4>(lut, fx0 + 1.5f, fy0 + 1.5f);
        cov[0] = fmaf(t11.x, k11, fmaf(t01.x, k01, fmaf(t10.x, k10, fmaf(t00.x, k00, cov[0]))));
        cov[1] = fmaf(t11.y, k11, fmaf(t01.y, k01, fmaf(t10.y, k10, fmaf(t00.y, k00, cov[1]))));
        cov[2] = fmaf(t11.z, k11, fmaf(t01.z, k01, fmaf(t10.z, k10, fmaf(t00.z, k00, cov[2]))));
        cov[3] = fmaf(t11.w, k11, fmaf(t01.w, k01, fmaf(t10.w, k10, fmaf(t00.w, k00, cov[3]))));
    } else if (kind == 1) {
        cov[0] += dX; cov[1] += dX; cov[2] += dX; cov[3] += dX;
    }
}

// Adds the signed area coverage of fills [begin, end) to the lane's 8 pixels (computeCoverage, fill.comp:51-71,
// for the lane's two 4-row groups: rows h*4 .. h*4+3 and 8+h*4 .. 8+h*4+3, column c). What only depends on the fill
// (left end point, slope) is computed once by the lane that loaded it and broadcast by shuffle.
__device__ __forceinline__ void accumulate_fills(const uint2 *__restrict__ fills, uint32_t begin, uint32_t end,
                                                 unsigned lane, cudaTextureObject_t lut, bool band, float cov[8]) {
    const float col = (float)(lane & 15u);  // left edge of the pixel column; tileFragCoord.x = col + 0.5
    const float fragy0 = (float)((lane >> 4) * 4u) + 0.5f;
    for (uint32_t at = begin; at < end; at += 32) {
        float x_from = 0.f, x_to = 0.f, ly = 0.f, d = 0.f;
        if (at + lane < end) {
            const uint2 f = __ldg(&fills[at + lane]);
            x_from = (float)(f.x & 0xffffu) * (1.0f / 256.0f);
            x_to = (float)(f.y & 0xffffu) * (1.0f / 256.0f);
            const float y_from = (float)(f.x >> 16) * (1.0f / 256.0f), y_to = (float)(f.y >> 16) * (1.0f / 256.0f);
            const bool from_left = x_from < x_to;  // fill.comp:53-55
            ly = from_left ? y_from : y_to;
            const float ry = from_left ? y_to : y_from;
            d = (ry - ly) * __fdividef(1.0f, fabsf(x_to - x_from));  // fill.comp:64 (bin never emits x_from == x_to)
        }
        const int m = (int)min(32u, end - at);
        for (int k = 0; k < m; k++) {
            const float xf = __shfl_sync(0xffffffffu, x_from, k), xt = __shfl_sync(0xffffffffu, x_to, k);
            // window = clamp(vec2(from.x, to.x), -0.5, 0.5) in fragment-centred coordinates (fill.comp:58) == saturate
            // in column-edge coordinates; the differences are exact (multiples of 1/256 below 16)
            const float wx = __saturatef(xf - col), wy = __saturatef(xt - col);
            const float lyk = __shfl_sync(0xffffffffu, ly, k), dk = __shfl_sync(0xffffffffu, d, k);
            const float dX = wx - wy;
            if (dX == 0.0f) continue;  // the fill does not overlap this pixel column: contributes exactly 0
            // y of the line at the middle of the window (fill.comp:59-63), relative to the group's first pixel centre
            const float offset = fmaf(0.5f, wx + wy, col - fminf(xf, xt));
            const float y = fmaf(dk, offset, lyk) - fragy0;
            // LUT texel coordinates: u = (y + 8) / 16, v = |d * dX| / 16 on 256 x 256 texels, minus the half texel
            const float lut_x = fmaf(y, 16.0f, 127.5f), lut_y = fmaf(fabsf(dk * dX), 16.0f, -0.5f);
            const float fx0 = floorf(lut_x), fy0 = floorf(lut_y);
            const float ax = lut_x - fx0, ay = lut_y - fy0;
            const float w11 = ax * ay, w10 = ax - w11, w01 = ay - w11, w00 = (1.0f - ax) - w01;
            const float half = fmaf(0.5f, lut_y, 1.0f);
            add_group(lut, band, lut_x, fx0, fy0, half, w00 * dX, w10 * dX, w01 * dX, w11 * dX, dX, cov);
            add_group(lut, band, lut_x - 128.0f, fx0 - 128.0f, fy0, half, w00 * dX, w10 * dX, w01 * dX, w11 * dX, dX,
                      cov + 4);  // 8 rows further down
        }
    }
}

// round(v * 255) for v in [0, 1] in the low byte (adding 1.5 * 2^23 leaves the integer, nearest-even, in the mantissa)
__device__ __forceinline__ uint32_t unorm8_bits(float v) { return __float_as_uint(fmaf(v, 255.0f, 12582912.0f)); }

// Coverage -> the 8 mask bytes of the lane (fill.comp:131-153): fill rule, optional min() with the clip mask,
// RGBA8-unorm quantisation.
__device__ __forceinline__ uint2 quantise_mask(const float cov[8], bool winding, const uint8_t *masks, int clip_alpha,
                                               unsigned lane) {
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; q++) {  // fill.comp:133-140
        if (winding) v[q] = fminf(fabsf(cov[q]), 1.0f);
        else v[q] = __saturatef(1.0f - fabsf(1.0f - glsl_mod(cov[q], 2.0f)));
    }
    if (clip_alpha >= 0) {  // fill.comp:147-150
        const uint2 clip = __ldg(reinterpret_cast<const uint2 *>(masks + (size_t)clip_alpha * 256) + lane);
#pragma unroll
        for (int q = 0; q < 8; q++)
            v[q] = fminf(v[q], (float)(((q < 4 ? clip.x : clip.y) >> (8 * (q & 3))) & 0xffu) * (1.0f / 255.0f));
    }
    uint2 out;
    out.x = __byte_perm(__byte_perm(unorm8_bits(v[0]), unorm8_bits(v[1]), 0x0040),
                        __byte_perm(unorm8_bits(v[2]), unorm8_bits(v[3]), 0x0040), 0x5410);
    out.y = __byte_perm(__byte_perm(unorm8_bits(v[4]), unorm8_bits(v[5]), 0x0040),
                        __byte_perm(unorm8_bits(v[6]), unorm8_bits(v[7]), 0x0040), 0x5410);
    return out;
}

constexpr int FILL_WARPS = 8;
constexpr float FILL_SCALE = 1048576.0f;  // coverage is accumulated in 12.20 fixed point (order-independent sums)
constexpr int FILL_ONE = 1 << 20;

// Standalone fill kernel: one warp per alpha tile, and inside the tile one lane per (fill, pixel column) PAIR.
//
// A fill only touches the pixel columns it spans (5 of 16 on tiger 4096^2) and, in each of them, the few rows around the
// line; below those rows its contribution is the constant dX, above them 0 (the area LUT saturates there, which
// pfcu_set_area_lut verifies for the uploaded LUT). So the work of a tile is enumerated as pairs -- a prefix sum over
// the fills' column spans, lane p takes pair p -- instead of giving every lane a fixed pixel column and every fill to
// every lane (30 % useful lanes). A pair samples only the 4-row groups its line passes through (one or two unless the
// line is steep) and adds what it finds to a 16 x 16 accumulator in shared memory as DIFFERENCES down
// its column: the tile's coverage is then one prefix sum per column, and "every row below gets dX" is a single add.
// Sums are fixed point, so the result does not depend on the order in which pairs or atomics land.
struct FillShared {
    int acc[FILL_WARPS][17][16];  // [row][column]; row 16 is a sink for differences that fall below the tile
};

__global__ void __launch_bounds__(FILL_WARPS * 32) k_fill(BatchView b, PaintView p) {
    __shared__ FillShared sh;
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t first_alpha = b.counters->first_alpha;
    uint32_t n_alpha = b.counters->n_alpha;
    if (n_alpha > b.alpha_capacity) n_alpha = b.alpha_capacity;
    int *const acc = &sh.acc[wib][0][0];
    for (int i = (int)lane; i < 17 * 16; i += 32) acc[i] = 0;
    __syncwarp();
    const bool band = p.lut_band != 0;
    const int c_own = (int)(lane & 15u), h_own = (int)(lane >> 4);

    for (uint32_t a = warp; a < n_alpha; a += n_warps) {
        const uint32_t id = first_alpha + a;
        if (id >= b.mask_capacity) break;
        // tile | winding << 31, clip mask slot, first fill, backdrop | fill count << 8
        const uint4 at = __ldg(reinterpret_cast<const uint4 *>(&b.alpha_tiles[a]));
        if ((at.x & 0x7fffffffu) >= b.tile_count) continue;  // fills that the clip made invisible: no mask needed
        const uint32_t count = at.w >> 8;
        uint32_t begin = at.z;
        if (begin > b.fill_capacity) begin = b.fill_capacity;
        const uint32_t end = min(begin + count, b.fill_capacity);

        for (uint32_t chunk = begin; chunk < end; chunk += 32) {
            // ---- one fill per lane: what only depends on the fill (fill.comp:53-64)
            const bool vf = chunk + lane < end;
            float x_from = 0.f, x_to = 0.f, ly = 0.f, d = 0.f;
            int c0 = 0, len = 0;
            if (vf) {
                const uint2 f = __ldg(&b.fills[chunk + lane]);
                x_from = (float)(f.x & 0xffffu) * (1.0f / 256.0f);
                x_to = (float)(f.y & 0xffffu) * (1.0f / 256.0f);
                const float y_from = (float)(f.x >> 16) * (1.0f / 256.0f), y_to = (float)(f.y >> 16) * (1.0f / 256.0f);
                const bool from_left = x_from < x_to;
                ly = from_left ? y_from : y_to;
                const float ry = from_left ? y_to : y_from;
                d = (ry - ly) * __fdividef(1.0f, fabsf(x_to - x_from));  // bin never emits x_from == x_to
                // pixel columns whose window [c, c + 1] the fill overlaps with positive length
                const float xmin = fminf(x_from, x_to), xmax = fmaxf(x_from, x_to);
                c0 = min((int)xmin, 15);
                const int c1 = min((int)ceilf(xmax) - 1, 15);
                len = max(c1 - c0 + 1, 1);
            }
            const int incl = (int)warp_incl_scan_u32((uint32_t)len, lane);
            const int excl = incl - len;
            const int n_pairs = __shfl_sync(0xffffffffu, incl, 31);
            const int pair_base = excl - c0;  // column of pair p of this fill = p - pair_base

            for (int p0 = 0; p0 < n_pairs; p0 += 32) {
                // ---- which fill does pair p0 + lane belong to: the number of fills that start at or before it, minus one
                // (every fill owns at least one pair, so fill k is lane k)
                const unsigned starts = __reduce_or_sync(0xffffffffu, (vf && excl >= p0 && excl < p0 + 32) ? 1u << (excl - p0) : 0u);
                const int before = __popc(__ballot_sync(0xffffffffu, vf && excl < p0));
                const int k = before - 1 + __popc(starts & (0xffffffffu >> (31 - lane)));
                const int srcl = k < 0 ? 0 : (k > 31 ? 31 : k);
                const float xf = __shfl_sync(0xffffffffu, x_from, srcl), xt = __shfl_sync(0xffffffffu, x_to, srcl);
                const float lyk = __shfl_sync(0xffffffffu, ly, srcl), dk = __shfl_sync(0xffffffffu, d, srcl);
                const int pb = __shfl_sync(0xffffffffu, pair_base, srcl);
                const int pidx = p0 + (int)lane;
                if (pidx >= n_pairs) continue;
                const int c = pidx - pb;
                const float col = (float)c;
                // window = clamp(vec2(from.x, to.x), -0.5, 0.5) in fragment-centred coordinates (fill.comp:58)
                const float wx = __saturatef(xf - col), wy = __saturatef(xt - col);
                const float dX = wx - wy;
                if (dX == 0.0f) continue;
                // y of the line at the middle of the window (fill.comp:59-63), tile space
                const float y_line = fmaf(dk, fmaf(0.5f, wx + wy, col - fminf(xf, xt)), lyk);
                const float lut_y = fmaf(fabsf(dk * dX), 16.0f, -0.5f);  // v * 256 - 0.5
                const float fy0 = floorf(lut_y), ay = lut_y - fy0;
                // rows whose LUT value is not saturated: above r_lo the coverage is 0, below r_hi it is 1
                int r_lo = 0, r_hi = 15;
                if (band) {
                    const float hw = (lut_y + 1.0f) * (1.0f / 32.0f);
                    r_lo = (int)ceilf(y_line - hw - (17.0f / 16.0f + 1.0f / 64.0f));
                    r_hi = (int)floorf(y_line + hw + 1.0f / 64.0f);
                }
                // windows are the LUT's own 4-row groups (rows 0-3, 4-7, ...): its channels are NOT exact one-row shifts of
                // each other for steep lines, so a row must be read from the channel fill.comp reads it from
                const int r_start = max(r_lo, 0) & ~3, r_end = min(r_hi, 15);
                const float ks = dX * FILL_SCALE;
                const int v_full = __float2int_rn(ks);
                int prev = 0;
                int *const colp = acc + c;
                int r0 = r_start;
                for (; r0 <= r_end; r0 += 4) {
                    // texture(uAreaLUT, vec2((y + 8) / 16, v)) for the 4 rows r0 .. r0 + 3 (fill.comp:66-70), fp32 weights
                    const float lut_x = fmaf(y_line - (float)r0, 16.0f, 119.5f);
                    const float fx0 = floorf(lut_x), ax = lut_x - fx0;
                    const float w11 = ax * ay, w10 = ax - w11, w01 = ay - w11, w00 = (1.0f - ax) - w01;
                    const float k00 = w00 * ks, k10 = w10 * ks, k01 = w01 * ks, k11 = w11 * ks;
                    const float4 t00 = tex2D<float4>(p.lut_tex, fx0 + 0.5f, fy0 + 0.5f), t10 = tex2D<float4>(p.lut_tex, fx0 + 1.5f, fy0 + 0.5f);
                    const float4 t01 = tex2D<float4>(p.lut_tex, fx0 + 0.5f, fy0 + 1.5f), t11 = tex2D<float4>(p.lut_tex, fx0 + 1.5f, fy0 + 1.5f);
                    const int v0 = __float2int_rn(fmaf(t11.x, k11, fmaf(t01.x, k01, fmaf(t10.x, k10, t00.x * k00))));
                    const int v1 = __float2int_rn(fmaf(t11.y, k11, fmaf(t01.y, k01, fmaf(t10.y, k10, t00.y * k00))));
                    const int v2 = __float2int_rn(fmaf(t11.z, k11, fmaf(t01.z, k01, fmaf(t10.z, k10, t00.z * k00))));
                    const int v3 = __float2int_rn(fmaf(t11.w, k11, fmaf(t01.w, k01, fmaf(t10.w, k10, t00.w * k00))));
                    atomicAdd(colp + r0 * 16, v0 - prev);  // rows past 15 land in the sink row or beyond? no: r0 <= 15
                    atomicAdd(colp + min(r0 + 1, 16) * 16, v1 - v0);
                    atomicAdd(colp + min(r0 + 2, 16) * 16, v2 - v1);
                    atomicAdd(colp + min(r0 + 3, 16) * 16, v3 - v2);
                    prev = v3;
                }
                // every row below the sampled ones is fully covered by the window: one add (telescopes with `prev`)
                const int tail = r0 > r_start ? r0 : max(r_hi + 1, 0);
                if (tail <= 15) atomicAdd(colp + tail * 16, v_full - prev);
            }
        }
        __syncwarp();
        // ---- coverage = prefix sum down each column (+ backdrop), fill rule, clip, RGBA8-unorm quantisation
        // (fill.comp:131-153); lane (c, h) owns rows h*4 .. h*4+3 and 8+h*4 .. 8+h*4+3 of column c
        int run = (int)(int8_t)(at.w & 0xffu) * FILL_ONE;
        int cv[8];
#pragma unroll
        for (int r = 0; r < 16; r++) {
            run += acc[r * 16 + c_own];
            const int q = (r & 3) + ((r >> 3) << 2);         // slot of row r for the lane that owns it
            if (((r >> 2) & 1) == h_own) cv[q] = run;        // rows 0-3, 8-11 -> h 0; rows 4-7, 12-15 -> h 1
        }
        __syncwarp();
        for (int i = (int)lane; i < 17 * 16; i += 32) acc[i] = 0;
        const bool winding = (at.x >> 31) != 0;
        uint2 clip = make_uint2(0xffffffffu, 0xffffffffu);
        if ((int)at.y >= 0 && at.y < b.mask_capacity) clip = __ldg(reinterpret_cast<const uint2 *>(b.masks + (size_t)at.y * 256) + lane);
        uint32_t bytes[8];
#pragma unroll
        for (int q = 0; q < 8; q++) {
            int v = cv[q];
            if (winding) v = min(abs(v), FILL_ONE);
            else { v &= 2 * FILL_ONE - 1; v = FILL_ONE - abs(FILL_ONE - v); }  // 1 - |1 - mod(cv, 2)|
            const uint32_t byte = ((uint32_t)v * 255u + (1u << 19)) >> 20;     // round(v * 255)
            bytes[q] = min(byte, ((q < 4 ? clip.x : clip.y) >> (8 * (q & 3))) & 0xffu);  // fill.comp:147-150
        }
        uint2 m;
        m.x = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
        m.y = bytes[4] | (bytes[5] << 8) | (bytes[6] << 16) | (bytes[7] << 24);
        reinterpret_cast<uint2 *>(b.masks + (size_t)id * 256)[lane] = m;
        __syncwarp();
    }
}

cudaError_t launch_fill(const BatchView &b, const PaintView &p, cudaStream_t s) {
    if (!b.tile_count || !p.lut_tex) return cudaSuccess;
    k_fill<<<sm_count() * 8, FILL_WARPS * 32, 0, s>>>(b, p);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------ tile

struct ColorSampler {
    const uint8_t *px;
    int w, h;
    bool repeat_u, repeat_v, nearest;
    __device__ __forceinline__ float4 operator()(float u, float v) const {
        return sample_rgba8(px, w, h, u, v, repeat_u, repeat_v, nearest);
    }
};

// filterRadialGradient, tile.comp:319-347
__device__ float4 filter_radial(const ColorSampler &cs, float cu, float cv, float4 p0, float4 p1) {
    const float dpx = cu - p0.x, dpy = cv - p0.y, dcx = p0.z, dcy = p0.w, dr = p1.y - p1.x;
    const float a = (dcx * dcx + dcy * dcy) - dr * dr;
    const float bq = (dpx * dcx + dpy * dcy) + p1.x * dr;
    const float c = (dpx * dpx + dpy * dpy) - p1.x * p1.x;
    const float discrim = bq * bq - a * c;
    float4 color = make_float4(0.f, 0.f, 0.f, 0.f);
    if (discrim != 0.0f) {
        const float sq = sqrtf(discrim);
        float tsx = (sq + bq) / a, tsy = (-sq + bq) / a;
        if (tsx > tsy) {
            const float tmp = tsx;
            tsx = tsy;
            tsy = tmp;
        }
        const float t = tsx >= 0.0f ? tsx : tsy;
        color = cs(p1.z + t, p1.w);
    }
    return color;
}

// filterBlur, tile.comp:354-392
__device__ float4 filter_blur(const ColorSampler &cs, float cu, float cv, float4 p0, float4 p1) {
    const float sox = p0.x / (float)cs.w, soy = p0.y / (float)cs.h;
    const int support = (int)p0.z;
    float gx = p1.x, gy = p1.y;
    const float gz = p1.z;
    float gauss_sum = gx;
    float4 color = cs(cu, cv);
    color.x *= gx; color.y *= gx; color.z *= gx; color.w *= gx;
    gx *= gy;
    gy *= gz;
    for (int i = 1; i <= support; i += 2) {
        float partial = gx;
        gx *= gy;
        gy *= gz;
        partial += gx;
        const float k = (float)i + gx / partial;
        const float4 a = cs(cu - sox * k, cv - soy * k), bq = cs(cu + sox * k, cv + soy * k);
        color.x += (a.x + bq.x) * partial;
        color.y += (a.y + bq.y) * partial;
        color.z += (a.z + bq.z) * partial;
        color.w += (a.w + bq.w) * partial;
        gauss_sum += 2.0f * partial;
        gx *= gy;
        gy *= gz;
    }
    color.x /= gauss_sum; color.y /= gauss_sum; color.z /= gauss_sum; color.w /= gauss_sum;
    return color;
}

// composite helpers, tile.comp:459-562
__device__ __forceinline__ float comp_div(float n, float d) { return d != 0.0f ? n / d : 0.0f; }
__device__ void rgb_to_hsl(const float rgb[3], float hsl[3]) {
    const float v = fmaxf(fmaxf(rgb[0], rgb[1]), rgb[2]), xmin = fminf(fminf(rgb[0], rgb[1]), rgb[2]);
    const float c = v - xmin, l = mixf(xmin, v, 0.5f);
    float t0, t1, t2;
    if (rgb[0] == v) { t0 = 0.0f; t1 = rgb[1]; t2 = rgb[2]; }
    else if (rgb[1] == v) { t0 = 2.0f; t1 = rgb[2]; t2 = rgb[0]; }
    else { t0 = 4.0f; t1 = rgb[0]; t2 = rgb[1]; }
    hsl[0] = 1.0471975511965976f * comp_div(t0 * c + t1 - t2, c);
    hsl[1] = comp_div(c, v);
    hsl[2] = l;
}
__device__ void hsl_to_rgb(const float hsl[3], float rgb[3]) {
    const float a = hsl[1] * fminf(hsl[2], 1.0f - hsl[2]);
    const float off[3] = {0.0f, 8.0f, 4.0f};
    for (int i = 0; i < 3; i++) {
        const float ks = glsl_mod(off[i] + hsl[0] * 1.9098593171027443f, 12.0f);
        rgb[i] = hsl[2] - clampf(fminf(ks - 3.0f, 9.0f - ks), -1.0f, 1.0f) * a;
    }
}
__device__ __forceinline__ float screen1(float d, float s) { return d + s - d * s; }
__device__ __forceinline__ float hard_light1(float d, float s) {
    return s <= 0.5f ? d * 2.0f * s : screen1(d, 2.0f * s - 1.0f);
}
__device__ __forceinline__ float color_dodge1(float d, float s) {
    return d == 0.0f ? 0.0f : (s == 1.0f ? 1.0f : d / (1.0f - s));
}
__device__ __forceinline__ float soft_light1(float d, float s) {
    const float dark = d <= 0.25f ? ((16.0f * d - 12.0f) * d + 4.0f) * d : sqrtf(d);
    const float factor = s <= 0.5f ? d * (1.0f - d) : dark - d;
    return d + (s * 2.0f - 1.0f) * factor;
}
__device__ void composite_rgb(const float d[3], const float s[3], int op, float out[3]) {
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
        const float dd = d[i], ss = s[i];
        float r;
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

// calculateColor, tile.comp:611-675: premultiplied source colour of one layer at one pixel.
template <bool SOLID>
__device__ __forceinline__ float4 shade(const Paint &pc, const ColorSampler &cs, float fragx, float fragy,
                                        float mask_alpha, float fb_w, float fb_h) {
    float4 color = pc.base;
    if (!SOLID) {
        const int combine = (pc.ctrl >> 8) & 0x3;
        if (combine != 0) {
            const float cu = pc.m0.x * fragx + pc.m0.z * fragy + pc.m1.x;
            const float cv = pc.m0.y * fragx + pc.m0.w * fragy + pc.m1.y;
            const int filter = (pc.ctrl >> 4) & 0xf;
            float4 c0;
            if (filter == 0x1) c0 = filter_radial(cs, cu, cv, pc.fp0, pc.fp1);
            else if (filter == 0x3) c0 = filter_blur(cs, cu, cv, pc.fp0, pc.fp1);
            else c0 = cs(cu, cv);
            if (combine == 0x1) color = make_float4(c0.x, c0.y, c0.z, c0.w * color.w);  // SRC_IN, tile.comp:128-129
            else if (combine == 0x2) color.w = c0.w * color.w;                          // DEST_IN, tile.comp:130-131
        }
    }
    color.w *= mask_alpha;
    if (!SOLID) {
        const int op = (pc.ctrl >> 10) & 0xf;
        if (op != 0) {  // composite(), tile.comp:564-582; its "dest" is the colour texture (FIXME upstream, :820-826)
            const float4 dc = cs(fragx / fb_w, fragy / fb_h);
            const float d[3] = {dc.x, dc.y, dc.z}, s[3] = {color.x, color.y, color.z};
            float blended[3];
            composite_rgb(d, s, op, blended);
            const float sa = color.w, da = dc.w;
            color.x = sa * (1.0f - da) * color.x + sa * da * blended[0] + (1.0f - sa) * dc.x;
            color.y = sa * (1.0f - da) * color.y + sa * da * blended[1] + (1.0f - sa) * dc.y;
            color.z = sa * (1.0f - da) * color.z + sa * da * blended[2] + (1.0f - sa) * dc.z;
            color.w = 1.0f;
        }
    }
    color.x *= color.w;
    color.y *= color.w;
    color.z *= color.w;
    return color;
}

constexpr int CT_WARPS = 8;    // warps per CTA
#ifndef CT_MIN_CTAS
#define CT_MIN_CTAS 3
#endif
constexpr int CT_TILES = 32;   // consecutive framebuffer tiles a CTA stages and renders
constexpr int CT_PRIMS = 160;  // list entries staged in shared memory (longer batches read their lists from global)

__device__ __forceinline__ uint32_t pack_rgba8(float4 c) {
    return __byte_perm(__byte_perm(unorm8_bits(__saturatef(c.x)), unorm8_bits(__saturatef(c.y)), 0x0040),
                       __byte_perm(unorm8_bits(__saturatef(c.z)), unorm8_bits(__saturatef(c.w)), 0x0040), 0x5410);
}

__device__ __forceinline__ float4 unpack_rgba8(uint32_t v) {
    const float k = 1.0f / 255.0f;
    return make_float4((float)(v & 0xffu) * k, (float)((v >> 8) & 0xffu) * k, (float)((v >> 16) & 0xffu) * k,
                       (float)(v >> 24) * k);
}

__device__ __forceinline__ void blend_over(float4 &dest, const float4 &src) {  // tile.comp:841
    const float ia = 1.0f - src.w;
    dest.x = fmaf(dest.x, ia, src.x);
    dest.y = fmaf(dest.y, ia, src.y);
    dest.z = fmaf(dest.z, ia, src.z);
    dest.w = fmaf(dest.w, ia, src.w);
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

constexpr int CT_SLOTS = 4;  // coverage masks a warp keeps in shared memory between its two phases

struct CompositeShared {
    uint4 fb[CT_TILES];                     // begin, count, z, cursor of the CTA's tiles
    uint4 prims[CT_PRIMS][2];               // their lists (contiguous in memory because the list offsets come from a scan)
    uint2 cover[CT_WARPS][CT_SLOTS][32];    // fused: mask bytes of the layers being blended, lane-major
    uint32_t flat_color[CT_TILES];          // packed RGBA8 of the tiles that are one colour
    uint8_t work[CT_TILES];                 // the other tiles (indices), compacted
    uint32_t n_work;
    uint32_t next;                          // dynamic tile distribution inside the CTA
};

// Walks a tile's list in paint order. Lists of up to 32 entries live in registers (one entry per lane, ranked by
// key); longer ones are walked by repeated selection of the next larger key from the list itself.
struct ListCursor {
    uint32_t layer, last_key;
    bool first;
};

template <bool FUSED>
__device__ __forceinline__ bool next_entry(ListCursor &cur, bool in_regs, uint32_t n, uint32_t n_sorted, int z,
                                           const uint4 *list, const uint4 &e0, const uint4 &e1, uint32_t rank,
                                           unsigned lane, uint4 &q0, uint4 &q1) {
    q1 = make_uint4(0u, 0xffffffffu, 0u, 0u);
    if (in_regs) {
        if (cur.layer >= n_sorted) return false;
        const int src = __ffs(__ballot_sync(0xffffffffu, rank == cur.layer)) - 1;
        q0.x = __shfl_sync(0xffffffffu, e0.x, src);
        q0.y = __shfl_sync(0xffffffffu, e0.y, src);
        q0.z = __shfl_sync(0xffffffffu, e0.z, src);
        q0.w = 0u;
        if (FUSED) {
            q0.w = __shfl_sync(0xffffffffu, e0.w, src);
            q1.x = __shfl_sync(0xffffffffu, e1.x, src);
            q1.y = __shfl_sync(0xffffffffu, e1.y, src);
            q1.z = __shfl_sync(0xffffffffu, e1.z, src);
        }
        cur.layer++;
        return true;
    }
    // selection: next smallest key >= z that is greater than the last one processed
    uint32_t best = 0xffffffffu, best_i = 0;
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t key = list[i * 2].x;
        if ((int)key >= z && (cur.first || key > cur.last_key) && key < best) {
            best = key;
            best_i = i;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob < best) {
            best = ob;
            best_i = oi;
        }
    }
    if (best == 0xffffffffu) return false;
    q0 = list[best_i * 2];
    if (FUSED) q1 = list[best_i * 2 + 1];
    cur.last_key = best;
    cur.first = false;
    cur.layer++;
    return true;
}

// One CTA renders CT_TILES consecutive framebuffer tiles. The three dependent loads of a tile (its list header, its
// list, the fills / masks the list points to) are issued for ALL of the CTA's tiles at once -- header and lists with
// coalesced loads into shared memory, fills / masks as L1 prefetches -- so their latency is paid once per CTA instead
// of once per tile; warps then pull tiles from a shared counter, which balances tiles with deep lists against empty
// ones. A tile is rendered in two phases so that the registers of the coverage computation and of the 8 x RGBA
// destination pixels are never live together: first the masks of its layers (fused mode), then the blend.
template <bool SOLID, bool FUSED>
__global__ void __launch_bounds__(CT_WARPS * 32, CT_MIN_CTAS) k_composite(BatchView b, PaintView p, TargetView tg,
                                                                          int clear, float4 clear_color, int origin) {
    __shared__ CompositeShared sh;
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t n_fb = (uint32_t)(b.fb_tw * b.fb_th);
    const uint32_t map0 = blockIdx.x * CT_TILES;
    const uint32_t n_tiles = min((uint32_t)CT_TILES, n_fb - map0);

    // ---- stage 1: list headers
    if (threadIdx.x < CT_TILES) {
        uint4 f = make_uint4(0u, 0u, 0u, 0u);
        if (threadIdx.x < n_tiles) f = __ldg(reinterpret_cast<const uint4 *>(&b.fb[map0 + threadIdx.x]));
        sh.fb[threadIdx.x] = f;
    }
    if (threadIdx.x == 0) sh.next = 0;
    __syncthreads();
    // ---- stage 2: the lists of all tiles are one contiguous range
    const uint32_t range0 = sh.fb[0].x;
    uint32_t range_n = sh.fb[n_tiles - 1].x + sh.fb[n_tiles - 1].y - range0;
    if (range0 + range_n > b.prim_capacity) range_n = range0 < b.prim_capacity ? b.prim_capacity - range0 : 0u;
    const bool staged = range_n <= CT_PRIMS;
    if (staged) {
        const uint4 *src = reinterpret_cast<const uint4 *>(b.prims + range0);
        for (uint32_t i = threadIdx.x; i < range_n * 2; i += CT_WARPS * 32) (&sh.prims[0][0])[i] = __ldg(src + i);
    }
    __syncthreads();
    // ---- stage 3: pull the fills (fused) or masks the lists point to towards this SM (warps 1..), while warp 0 sorts the
    // CTA's tiles into flat ones -- every surviving layer covers the whole tile with one colour, the common case for
    // interior and empty tiles: ONE THREAD blends them, as a single pixel -- and the rest.
    if (threadIdx.x >= 32) {
        if (staged) {
            for (uint32_t i = threadIdx.x - 32; i < range_n; i += (CT_WARPS - 1) * 32) {
                const uint4 q0 = sh.prims[i][0], q1 = sh.prims[i][1];
                if ((int)q0.y < 0) continue;
                if (FUSED && (q1.z & PRIM_OWNS_MASK)) {
                    if (q1.x && q0.w < b.fill_capacity) prefetch_l1(b.fills + q0.w);
                    if ((int)q1.y >= 0 && q1.y < b.mask_capacity) prefetch_l1(b.masks + (size_t)q1.y * 256);
                } else if (q0.y < b.mask_capacity) {
                    prefetch_l1(b.masks + (size_t)q0.y * 256);
                    prefetch_l1(b.masks + (size_t)q0.y * 256 + 128);
                }
            }
        }
    } else {
        const uint32_t t = threadIdx.x;
        bool is_flat = false, is_work = false;
        if (t < n_tiles) {
            const uint4 fbt = sh.fb[t];
            const uint32_t n = fbt.y;
            const int z = (int)fbt.z;
            if (n == 0) {
                is_flat = clear != 0;  // LOAD_ACTION_LOAD leaves an empty tile alone (tile.comp:743-744)
            } else if (!clear || !staged || n > 12) {
                is_work = true;
            } else {
                const uint4 *list = &sh.prims[fbt.x - range0][0];
                float4 dest = clear_color;
                uint32_t last_key = 0;
                bool first = true;
                is_flat = true;
                while (true) {  // next key in paint order (sort.comp:49-83), z-culled
                    uint32_t best = 0xffffffffu, best_i = 0;
                    for (uint32_t i = 0; i < n; i++) {
                        const uint32_t key = list[i * 2].x;
                        if ((int)key >= z && (first || key > last_key) && key < best) {
                            best = key;
                            best_i = i;
                        }
                    }
                    if (best == 0xffffffffu) break;
                    last_key = best;
                    first = false;
                    const uint4 q0 = list[best_i * 2];
                    const int tile_ctrl = (int)((q0.z >> 16) & 0xffu), backdrop = (int)q0.z >> 24;
                    if ((int)q0.y >= 0) {
                        if (tile_ctrl & 0x3) {  // a mask: per-pixel work
                            is_flat = false;
                            break;
                        }
                    } else if (backdrop != 0 && (tile_ctrl & 0x2) && (abs(backdrop) & 1) == 0) {
                        continue;  // tile.comp:786-792
                    }
                    const uint32_t color_entry = q0.z & 0xffffu;
                    float4 src = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (color_entry < p.n_paints) {
                        if (!SOLID && __ldg(&p.paints[color_entry].ctrl) != 0) {  // textured paint: per-pixel work
                            is_flat = false;
                            break;
                        }
                        src = __ldg(&p.paints[color_entry].base);
                    }
                    src.x *= src.w;
                    src.y *= src.w;
                    src.z *= src.w;
                    blend_over(dest, src);
                }
                is_work = !is_flat;
                if (is_flat) sh.flat_color[t] = pack_rgba8(dest);
            }
            if (n == 0 && is_flat) sh.flat_color[t] = pack_rgba8(clear_color);
        }
        const unsigned work_mask = __ballot_sync(0xffffffffu, is_work), flat_mask = __ballot_sync(0xffffffffu, is_flat);
        if (is_work) sh.work[__popc(work_mask & ((1u << t) - 1u))] = (uint8_t)t;
        if (t == 0) {
            sh.n_work = (uint32_t)__popc(work_mask);
            sh.fb[0].w = flat_mask;  // (the cursor word of the header is not needed any more)
        }
    }
    __syncthreads();

    // ---- stage 4a: flat tiles, stored by the whole CTA with 16-byte stores: consecutive threads write consecutive
    // 16-byte pieces of one pixel row across the CTA's tiles (2 KiB contiguous when the tiles share a tile row)
    {
        const uint32_t flat_mask = sh.fb[0].w;
        if (flat_mask) {
            for (uint32_t j = threadIdx.x; j < CT_TILES * TILE * 4; j += CT_WARPS * 32) {
                const uint32_t t = (j >> 2) & (CT_TILES - 1), row = j >> 7, quarter = j & 3u;
                if (!((flat_mask >> t) & 1u)) continue;
                const uint32_t map = map0 + t;
                const int tile_y = (int)(map / (uint32_t)b.fb_tw), tile_x = (int)map - tile_y * b.fb_tw;
                const int gy = tile_y * TILE + (int)row, gx = tile_x * TILE + (int)quarter * 4;
                if (gy >= tg.height) continue;
                const uint32_t px = sh.flat_color[t];
                uint8_t *dst = tg.pixels + (size_t)gy * tg.pitch + (size_t)gx * 4;
                if (gx + 3 < tg.width) {
                    *reinterpret_cast<uint4 *>(dst) = make_uint4(px, px, px, px);
                } else {
                    for (int k = 0; k < 4; k++)
                        if (gx + k < tg.width) reinterpret_cast<uint32_t *>(dst)[k] = px;
                }
            }
        }
    }

    ColorSampler cs;
    cs.px = p.color_px;
    cs.w = p.color_w;
    cs.h = p.color_h;
    cs.repeat_u = (p.sampling_flags & 1u) != 0;
    cs.repeat_v = (p.sampling_flags & 2u) != 0;
    cs.nearest = (p.sampling_flags & 0xcu) != 0;
    const int c = (int)(lane & 15u), h = (int)(lane >> 4);

    // ---- stage 4b: the other tiles, one warp per tile
    const uint32_t n_work = sh.n_work;
    while (true) {
        uint32_t wi = 0;
        if (lane == 0) wi = atomicAdd(&sh.next, 1u);
        wi = __shfl_sync(0xffffffffu, wi, 0);
        if (wi >= n_work) break;
        const uint32_t t = sh.work[wi];
        const uint32_t map = map0 + t;
        const uint4 fbt = sh.fb[t];
        uint32_t n = fbt.y;
        if (n == 0 && !clear) continue;  // tile.comp:743-744
        const uint32_t begin = fbt.x;
        if (begin + n > b.prim_capacity) n = begin < b.prim_capacity ? b.prim_capacity - begin : 0u;
        const int z = (int)fbt.z;
        const int tile_y = (int)(map / (uint32_t)b.fb_tw), tile_x = (int)map - tile_y * b.fb_tw;
        const int gx = tile_x * TILE + c, gy0 = tile_y * TILE + h * 4;  // pixel q of the lane: row gy0 + ROW(q)
#define ROW(q) ((q) + ((q) & 4))
        const bool interior = (tile_x + 1) * TILE <= tg.width && (tile_y + 1) * TILE <= tg.height;
        uint8_t *const px0 = tg.pixels + (size_t)gy0 * tg.pitch + (size_t)gx * 4;  // the lane's first pixel

        // Order by paint order and z-cull (sort.comp:49-83). Keys (dense tile indices) are unique.
        const bool in_regs = n <= 32;
        const uint4 *list = staged ? &sh.prims[begin - range0][0] : reinterpret_cast<const uint4 *>(b.prims + begin);
        uint4 e0 = make_uint4(0xffffffffu, 0xffffffffu, 0u, 0u), e1 = make_uint4(0u, 0xffffffffu, 0u, 0u);
        uint32_t rank = 0xffffffffu, n_sorted = 0;
        if (in_regs && n) {
            if (lane < n) {
                e0 = list[lane * 2];
                if (FUSED) e1 = list[lane * 2 + 1];
            }
            const bool keep = lane < n && (int)e0.x >= z;
            const uint32_t kept_mask = __ballot_sync(0xffffffffu, keep);
            n_sorted = (uint32_t)__popc(kept_mask);
            if (keep) {
                rank = 0;
                if (n_sorted > 1) {
                    for (uint32_t j = 0; j < n; j++) {  // keys of the other entries straight from the list
                        const uint32_t kj = list[j * 2].x;
                        rank += ((kept_mask >> j) & 1u) && kj < e0.x ? 1u : 0u;
                    }
                }
            }
        }

        // dest: one colour for the whole tile while `uniform`, else 8 pixels per lane
        bool uniform = clear != 0;
        float4 dest_u = clear_color;
        float4 dest[8];
        if (!uniform) {
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const bool ok = gx < tg.width && gy0 + ROW(q) < tg.height;
                dest[q] = ok ? unpack_rgba8(*reinterpret_cast<const uint32_t *>(px0 + (size_t)ROW(q) * tg.pitch))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }

        // gl_FragCoord of the full canvas (only textured paints look at it); render-target pages have no origin
        const float org_x = (float)(origin ? b.fb_tx0 * TILE : 0), org_y = (float)(origin ? b.fb_ty0 * TILE : 0);
        const float fragx = (float)gx + org_x + 0.5f;
        ListCursor cur = {0u, 0u, true};
        bool more = n != 0;
        while (more) {
            // ---- phase 1 (fused): coverage masks of the next layers that own one, up to CT_SLOTS of them
            ListCursor ahead = cur;
            uint32_t n_round = 0xffffffffu;  // layers of this round (all remaining ones unless the slots run out)
            if (FUSED) {
                int slot = 0;
                uint32_t seen = 0;
                uint4 q0, q1;
                while (next_entry<FUSED>(ahead, in_regs, n, n_sorted, z, list, e0, e1, rank, lane, q0, q1)) {
                    seen++;
                    const int tile_ctrl = (int)((q0.z >> 16) & 0xffu);
                    if ((int)q0.y < 0 || !(tile_ctrl & 0x3) || !(q1.z & PRIM_OWNS_MASK) || q0.y >= b.mask_capacity) continue;
                    float cov[8];
                    const float bd = (float)((int)q0.z >> 24);
#pragma unroll
                    for (int q = 0; q < 8; q++) cov[q] = bd;
                    uint32_t fb_ = q0.w, fe_ = q0.w + q1.x;
                    if (fe_ > b.fill_capacity) fe_ = b.fill_capacity;
                    if (fb_ > fe_) fb_ = fe_;
                    accumulate_fills(b.fills, fb_, fe_, lane, p.lut_tex, p.lut_band != 0, cov);
                    const int clip_alpha = (int)q1.y >= 0 && q1.y < b.mask_capacity ? (int)q1.y : -1;
                    sh.cover[wib][slot][lane] = quantise_mask(cov, (tile_ctrl & 0x1) != 0, b.masks, clip_alpha, lane);
                    if (++slot == CT_SLOTS) {
                        n_round = seen;
                        break;
                    }
                }
                __syncwarp();
            }
            // ---- phase 2: blend this round's layers
            int slot = 0;
            more = false;
            for (uint32_t i = 0; i < n_round; i++) {
                uint4 q0, q1;
                if (!next_entry<FUSED>(cur, in_regs, n, n_sorted, z, list, e0, e1, rank, lane, q0, q1)) break;
                more = i + 1 == n_round;  // stopped by the slot limit: another round follows
                // tile.comp:765-800
                const uint32_t color_entry = q0.z & 0xffffu;
                int tile_ctrl = (int)((q0.z >> 16) & 0xffu);
                const int backdrop = (int)q0.z >> 24;
                const int alpha = (int)q0.y;
                if (alpha < 0) {
                    if (backdrop != 0 && (tile_ctrl & 0x2) && (abs(backdrop) & 1) == 0) continue;  // tile.comp:786-792
                    tile_ctrl &= ~0x3;
                }
                const int mask_ctrl = tile_ctrl & 0x3;
                const bool masked = mask_ctrl != 0 && alpha >= 0 && (uint32_t)alpha < b.mask_capacity;

                Paint pc;
                if (color_entry < p.n_paints) {
                    pc.base = __ldg(&p.paints[color_entry].base);
                    if (!SOLID) {
                        pc.m0 = __ldg(&p.paints[color_entry].m0);
                        pc.m1 = __ldg(&p.paints[color_entry].m1);
                        pc.fp0 = __ldg(&p.paints[color_entry].fp0);
                        pc.fp1 = __ldg(&p.paints[color_entry].fp1);
                        pc.ctrl = __ldg(&p.paints[color_entry].ctrl);
                    }
                } else {
                    pc.base = pc.m0 = pc.m1 = pc.fp0 = pc.fp1 = make_float4(0.f, 0.f, 0.f, 0.f);
                    pc.ctrl = 0;
                }
                const bool flat_paint = SOLID || pc.ctrl == 0;

                if (!masked && flat_paint) {
                    // the whole tile gets one premultiplied colour (calculateColor with maskAlpha == 1, tile.comp:611-675)
                    float4 src = pc.base;
                    src.x *= src.w;
                    src.y *= src.w;
                    src.z *= src.w;
                    if (uniform) {
                        blend_over(dest_u, src);
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; q++) blend_over(dest[q], src);
                    }
                    continue;
                }
                if (uniform) {
#pragma unroll
                    for (int q = 0; q < 8; q++) dest[q] = dest_u;
                    uniform = false;
                }
                // the RGBA8 mask texel values tile.comp:594-598 would fetch for the lane's 8 pixels
                uint2 mask8 = make_uint2(0xffffffffu, 0xffffffffu);
                if (masked) {
                    if (FUSED && (q1.z & PRIM_OWNS_MASK)) mask8 = sh.cover[wib][slot++][lane];
                    else mask8 = __ldg(reinterpret_cast<const uint2 *>(b.masks + (size_t)alpha * 256) + lane);
                }
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    float mask_alpha = 1.0f;
                    if (masked) {  // sampleMask, tile.comp:586-607 (backdrop is 0 for alpha tiles)
                        float cov = (float)(((q < 4 ? mask8.x : mask8.y) >> (8 * (q & 3))) & 0xffu) * (1.0f / 255.0f);
                        if (!(mask_ctrl & 0x1)) cov = 1.0f - fabsf(1.0f - glsl_mod(cov, 2.0f));
                        mask_alpha = fminf(mask_alpha, cov);
                    }
                    float4 src;
                    if (flat_paint) {
                        src = pc.base;
                        src.w *= mask_alpha;
                        src.x *= src.w;
                        src.y *= src.w;
                        src.z *= src.w;
                    } else {
                        src = shade<false>(pc, cs, fragx, (float)(gy0 + ROW(q)) + org_y + 0.5f, mask_alpha, (float)tg.width,
                                           (float)tg.height);
                    }
                    blend_over(dest[q], src);
                }
            }
            if (FUSED) __syncwarp();  // the slots are rewritten by the next round
        }

        if (uniform) {
            const uint32_t px = pack_rgba8(dest_u);
            if (interior) {  // 2 x 16-byte stores per lane: row lane >> 1, pixels (lane & 1) * 8 .. + 7
                uint4 *d = reinterpret_cast<uint4 *>(tg.pixels + (size_t)(tile_y * TILE + (int)(lane >> 1)) * tg.pitch) +
                           (tile_x * 4 + (int)(lane & 1u) * 2);
                d[0] = make_uint4(px, px, px, px);
                d[1] = make_uint4(px, px, px, px);
            } else {
#pragma unroll
                for (int q = 0; q < 8; q++)
                    if (gx < tg.width && gy0 + ROW(q) < tg.height)
                        *reinterpret_cast<uint32_t *>(px0 + (size_t)ROW(q) * tg.pitch) = px;
            }
            continue;
        }
#pragma unroll
        for (int q = 0; q < 8; q++)
            if (interior || (gx < tg.width && gy0 + ROW(q) < tg.height))
                *reinterpret_cast<uint32_t *>(px0 + (size_t)ROW(q) * tg.pitch) = pack_rgba8(dest[q]);
    }
}

#undef ROW

cudaError_t launch_composite(const BatchView &b, const PaintView &p, const TargetView &t, int clear,
                             const float clear_color[4], int origin, cudaStream_t s) {
    if (b.fb_tw <= 0 || b.fb_th <= 0) return cudaSuccess;
    const uint32_t n_fb = (uint32_t)(b.fb_tw * b.fb_th);
    const unsigned grid = (n_fb + CT_TILES - 1) / CT_TILES;
    const float4 cc = make_float4(clear_color[0], clear_color[1], clear_color[2], clear_color[3]);
    const int threads = CT_WARPS * 32;
    if (p.fused) {
        if (p.all_solid) k_composite<true, true><<<grid, threads, 0, s>>>(b, p, t, clear, cc, origin);
        else k_composite<false, true><<<grid, threads, 0, s>>>(b, p, t, clear, cc, origin);
    } else {
        if (p.all_solid) k_composite<true, false><<<grid, threads, 0, s>>>(b, p, t, clear, cc, origin);
        else k_composite<false, false><<<grid, threads, 0, s>>>(b, p, t, clear, cc, origin);
    }
    return cudaGetLastError();
}

}  // namespace pfcu
