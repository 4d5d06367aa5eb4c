// fill + tile ("composite"): the two rasterizing stages, the HBM-bound end of the path.
//
// Pixel ownership is the same in both kernels: a WARP owns one 16 x 16 tile, lane (k = lane & 3, s = lane >> 2) owns
// pixel columns 4k .. 4k+3 of rows 2s and 2s+1 -- two 16-byte pieces of the framebuffer, one 8-byte piece of a mask.
//
//   fill : pathfinder/shaders/d3d11/fill.comp:51-154. Masks whose tile the z-buffer culls are skipped (nothing reads
//          them). The tile's fills are contiguous (CSR from the scatter), read
//          once with coalesced 8-byte loads. The work of a tile is enumerated as (fill, pixel column) pairs, one per
//          lane; a pair samples the 256 x 256 area LUT (behind the texture unit: texel fetch + unorm conversion, the
//          bilinear weights are applied in fp32) only for the 4-row groups its line passes through and adds the
//          result to a 16 x 16 fixed-point accumulator in shared memory as differences down its column.
//   tile : pathfinder/shaders/d3d11/tile.comp:737-850 with the shading functions of tile.comp:126-134 (combine),
//          :319-347 (radial gradient), :354-392 (blur), :459-582 (composite), :586-607 (mask), :694-726 (paint
//          metadata, decoded once per upload into a float table). A CTA stages the (z-culled) lists of 16 consecutive
//          framebuffer tiles in shared memory, orders them by paint order with 8 threads per tile (sort.comp:49-83)
//          and blends the leading whole-tile layers of every tile as ONE pixel. Tiles that end there are stored by
//          the whole CTA with 16-byte stores; the others go to one warp each, which blends the remaining layers in
//          registers with packed f32x2 arithmetic and stores 16 bytes per lane and row.
#include <cuda_fp16.h>

#include "pfcu_device.h"

namespace pfcu {

// ---- packed f32x2 arithmetic (sm_100a FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per issue slot, each half
// rounded exactly like the scalar instruction)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "mul.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
        "add.rn.f32x2 rd, ra, rb;\n\t"
        "mov.b64 {%0, %1}, rd;}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 splat(float v) { return make_float2(v, v); }

__device__ __forceinline__ float4 ld_rgba8(const uint8_t *px, int w, int x, int y) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(px) + (size_t)y * w + x);
    // float(byte) without the conversion pipe (a quarter-rate I2F per channel is what a blur tap would spend most of its
    // time on): the byte lands in the mantissa of 2^23 and the subtraction is exact, so the value is the same
    const float k = 1.0f / 255.0f, m = 8388608.0f;
    return make_float4((__uint_as_float(__byte_perm(v, 0x4b000000u, 0x7540)) - m) * k,
                       (__uint_as_float(__byte_perm(v, 0x4b000000u, 0x7541)) - m) * k,
                       (__uint_as_float(__byte_perm(v, 0x4b000000u, 0x7542)) - m) * k,
                       (__uint_as_float(__byte_perm(v, 0x4b000000u, 0x7543)) - m) * k);
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

// The same bilinear fetch with the two axes prepared separately (texel indices + weight per axis): a blur along x keeps the y
// half for all of its taps and the other way round. Same operations in the same order as sample_rgba8, so the same bits.
struct AxisTaps {
    int i0, i1;
    float a;
};
__device__ __forceinline__ AxisTaps axis_taps(float coord, int n, bool repeat) {
    float x = coord * (float)n;
    x -= 0.5f;
    const float f0 = floorf(x);
    AxisTaps t;
    t.a = x - f0;
    t.i0 = wrap_or_clamp((int)f0, n, repeat);
    t.i1 = wrap_or_clamp((int)f0 + 1, n, repeat);
    return t;
}
// (packed f32x2: the byte -> float conversion and the three interpolations of two channels per instruction, each half
// rounded like the scalar instruction it replaces -- a + (b - a) * t is one subtraction and one fused multiply-add there too)
struct Texel2 {
    float2 rg, ba;
};
__device__ __forceinline__ Texel2 ld_rgba8_2(const uint8_t *px, int w, int x, int y) {
    const uint32_t v = __ldg(reinterpret_cast<const uint32_t *>(px) + (size_t)y * w + x);
    const float2 m = splat(-8388608.0f), k = splat(1.0f / 255.0f);
    Texel2 t;
    t.rg = mul2(add2(make_float2(__uint_as_float(__byte_perm(v, 0x4b000000u, 0x7540)), __uint_as_float(__byte_perm(v, 0x4b000000u, 0x7541))), m), k);
    t.ba = mul2(add2(make_float2(__uint_as_float(__byte_perm(v, 0x4b000000u, 0x7542)), __uint_as_float(__byte_perm(v, 0x4b000000u, 0x7543))), m), k);
    return t;
}
__device__ __forceinline__ float2 lerp2(float2 a, float2 b, float2 t) { return fma2(fma2(a, splat(-1.0f), b), t, a); }
__device__ __forceinline__ float4 sample_axes(const uint8_t *px, int w, const AxisTaps &X, const AxisTaps &Y) {
    const float2 ax = splat(X.a), ay = splat(Y.a);
    const Texel2 a = ld_rgba8_2(px, w, X.i0, Y.i0), b = ld_rgba8_2(px, w, X.i1, Y.i0);
    const Texel2 c = ld_rgba8_2(px, w, X.i0, Y.i1), d = ld_rgba8_2(px, w, X.i1, Y.i1);
    const float2 rg = lerp2(lerp2(a.rg, b.rg, ax), lerp2(c.rg, d.rg, ax), ay);
    const float2 ba = lerp2(lerp2(a.ba, b.ba, ax), lerp2(c.ba, d.ba, ax), ay);
    return make_float4(rg.x, rg.y, ba.x, ba.y);
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

// round(v * 255) for v in [0, 1] in the low byte (adding 1.5 * 2^23 leaves the integer, nearest-even, in the mantissa)
__device__ __forceinline__ uint32_t unorm8_bits(float v) { return __float_as_uint(fmaf(v, 255.0f, 12582912.0f)); }

constexpr int FILL_WARPS = 8;
#ifndef FILL_PERMUTE
#define FILL_PERMUTE 1
#endif
#ifndef FILL_CTAS_PER_SM
#define FILL_CTAS_PER_SM 4  // 64 registers x 256 threads: one wave
#endif
#ifndef FILL_GROUP_N
#define FILL_GROUP_N 4
#endif
constexpr int FILL_GROUP = FILL_GROUP_N;  // alpha tiles a warp rasterizes together (2 or 4)
constexpr float FILL_SCALE = 1048576.0f;  // coverage is accumulated in 12.20 fixed point (order-independent sums)
constexpr int FILL_ONE = 1 << 20;
constexpr int FILL_CS = 20;               // accumulator of one tile: [column][row], 16 rows padded to 20 words: a lane's
                                          // 16-byte loads of 8 consecutive rows are conflict free
constexpr int FILL_ACC = 16 * FILL_CS;

// Standalone fill kernel: a warp takes FILL_GROUP consecutive alpha tiles, and inside them one lane per
// (fill, pixel column) PAIR.
//
// A fill only touches the pixel columns it spans (5 of 16 on tiger 4096^2) and, in each of them, the few rows around the
// line; below those rows its contribution is the constant dX, above them 0 (the area LUT saturates there, which
// pfcu_set_area_lut verifies for the uploaded LUT). So the work is enumerated as pairs -- a prefix sum over the fills'
// column spans, lane p takes pair p -- instead of giving every lane a fixed pixel column and every fill to every lane
// (30 % useful lanes). A tile has 3 fills and 16 pairs on average, so the fills of a few consecutive tiles (contiguous:
// CSR) are taken together: 32 fills per load, full warps of pairs. A pair samples only the 4-row groups its line passes
// through (one or two unless the line is steep) and adds what it finds to its tile's 16 x 16 accumulator in shared
// memory as DIFFERENCES down its column: the tile's coverage is then one prefix sum per column, and "every row below
// gets dX" is a single add. Sums are fixed point, so the result does not depend on the order in which pairs or atomics
// land.
// what fill needs of PaintView
struct LutView {
    cudaTextureObject_t lut_tex;
    int lut_band;
};

struct __align__(16) FillShared {
    int acc[FILL_WARPS][FILL_GROUP][FILL_ACC];
};

// The masks of up to G alpha tiles, rasterized by one warp. Lane g < G holds the record of tile g in `at` (tile | winding
// << 31, clip mask slot, first fill, backdrop | fill count << 8; a tile index >= tile_count: nothing to do) and where its
// mask goes in `dst_code` (a mask slot of b.masks; with CACHE, bit 31 marks a 256-byte slot of the shared-memory array
// `cache` instead). `acc` = the warp's G accumulators: all zero on entry, all zero again on return.
template <int G, bool CACHE, class B, class P>
__device__ __forceinline__ void fill_group(uint4 at, const uint32_t dst_code, uint8_t *cache, int *const acc,
                                           const B &b, const P &p, const unsigned lane) {
    const bool band = p.lut_band != 0;
    const bool visible = (at.x & 0x7fffffffu) < b.tile_count;  // else: fills that the clip made invisible, no mask
    const uint32_t begin = min(at.z, b.fill_capacity);
    const uint32_t count = visible ? min(at.w >> 8, b.fill_capacity - begin) : 0u;
    // the fills of the group as one sequence: tile g owns [pre_g, pre_g + count_g)
    uint32_t incl = count;
#pragma unroll
    for (int d = 1; d < G; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= (unsigned)d) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, G - 1);
    uint32_t pre[G], rel[G];  // rel_g: address of the tile's fill j = rel_g + j
#pragma unroll
    for (int g = 0; g < G; g++) {
        pre[g] = __shfl_sync(0xffffffffu, incl - count, g);
        rel[g] = __shfl_sync(0xffffffffu, begin, g) - pre[g];
    }

    for (uint32_t chunk = 0; chunk < total; chunk += 32) {
        // ---- one fill per lane: what only depends on the fill (fill.comp:53-64)
        const uint32_t j = chunk + lane;
        const bool vf = j < total;
        float x_from = 0.f, x_to = 0.f, ly = 0.f, d = 0.f;
        int c0 = 0, len = 0, g_own = 0;
        if (vf) {
            uint32_t r = rel[0];
#pragma unroll
            for (int g = 1; g < G; g++)
                if (j >= pre[g]) {
                    g_own = g;
                    r = rel[g];
                }
            const uint2 f = __ldg(&b.fills[r + j]);
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
        const int incl_p = (int)warp_incl_scan_u32((uint32_t)len, lane);
        const int excl = incl_p - len;
        const int n_pairs = __shfl_sync(0xffffffffu, incl_p, 31);
        // column of pair p of this fill = p - pair_base; the fill's tile rides along in the low bits
        const int pair_base_g = ((excl - c0) << 2) | g_own;

        for (int p0 = 0; p0 < n_pairs; p0 += 32) {
            // ---- which fill does pair p0 + lane belong to: the number of fills that start at or before it, minus one
            // (every fill owns at least one pair, so fill k is lane k)
            const unsigned starts = __reduce_or_sync(0xffffffffu, (vf && excl >= p0 && excl < p0 + 32) ? 1u << (excl - p0) : 0u);
            const int before = __popc(__ballot_sync(0xffffffffu, vf && excl < p0));
            const int k = before - 1 + __popc(starts & (0xffffffffu >> (31 - lane)));
            const int srcl = k < 0 ? 0 : (k > 31 ? 31 : k);
            const float xf = __shfl_sync(0xffffffffu, x_from, srcl), xt = __shfl_sync(0xffffffffu, x_to, srcl);
            const float lyk = __shfl_sync(0xffffffffu, ly, srcl), dk = __shfl_sync(0xffffffffu, d, srcl);
            const int pbg = __shfl_sync(0xffffffffu, pair_base_g, srcl);
            const int pidx = p0 + (int)lane;
            if (pidx >= n_pairs) continue;
            const int c = pidx - (pbg >> 2);
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
            int *const colp = acc + (pbg & 3) * FILL_ACC + c * FILL_CS;
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
                atomicAdd(colp + r0, v0 - prev);  // r0 <= 12 (a multiple of 4): rows r0 .. r0 + 3 exist
                atomicAdd(colp + r0 + 1, v1 - v0);
                atomicAdd(colp + r0 + 2, v2 - v1);
                atomicAdd(colp + r0 + 3, v3 - v2);
                prev = v3;
            }
            // every row below the sampled ones is fully covered by the window: one add (telescopes with `prev`)
            const int tail = r0 > r_start ? r0 : max(r_hi + 1, 0);
            if (tail <= 15) atomicAdd(colp + tail, v_full - prev);
        }
    }
    __syncwarp();
    // ---- per tile: coverage = prefix sum down each column (+ backdrop), fill rule, clip, RGBA8-unorm quantisation
    // (fill.comp:131-153). Lane (c, h) = (lane & 15, lane >> 4) owns rows 8h .. 8h+7 of column c: two 16-byte loads of
    // the accumulator (zeroed for the next group at once: nobody else reads them), a serial prefix sum in registers and
    // ONE shuffle for the upper half's total (round 1 kept the mask's lane layout here -- 4 columns x 2 rows -- and paid
    // a 3-level shuffle scan of 4 values: 135 instructions per tile, 37 % of the kernel). The bytes then go to the
    // mask's layout by a 4 x 4 byte transpose among the 4 lanes of a column group (2 shuffles + 2 byte permutes per
    // word) and two 4-byte stores per lane.
    const int c_own = (int)(lane & 15u), h_own = (int)(lane >> 4), j_own = (int)(lane & 3u);
    // byte selectors of the two transpose rounds (lane-dependent, tile-independent)
    const uint32_t sel1 = (j_own & 1) ? 0x3715u : 0x6240u;  // odd: [t1, w1, t3, w3]; even: [w0, t0, w2, t2]
    const uint32_t sel2 = (j_own & 2) ? 0x3276u : 0x5410u;  // upper: [u2, u3, w2, w3]; lower: [w0, w1, u0, u1]
    // after the transpose this lane holds pixel row 8h + j (from the first word) and 8h + 4 + j (second word) of columns
    // 4k .. 4k+3; the mask's layout (pfcu_device.h) puts row y, columns 4k .. 4k+3 at byte ((y >> 1) * 4 + k) * 8 + (y & 1) * 4
    const int k_grp = c_own >> 2, y_a = h_own * 8 + j_own, y_b = y_a + 4;
    const uint32_t off_a = (uint32_t)(((y_a >> 1) * 4 + k_grp) * 8 + (y_a & 1) * 4);
    const uint32_t off_b = (uint32_t)(((y_b >> 1) * 4 + k_grp) * 8 + (y_b & 1) * 4);
#pragma unroll
    for (int g = 0; g < G; g++) {
        const uint32_t at_x = __shfl_sync(0xffffffffu, at.x, g), at_y = __shfl_sync(0xffffffffu, at.y, g);
        const uint32_t at_w = __shfl_sync(0xffffffffu, at.w, g);
        const uint32_t code = __shfl_sync(0xffffffffu, dst_code, g);
        if ((at_x & 0x7fffffffu) >= b.tile_count) continue;  // (warp-uniform)
        int4 *const own = reinterpret_cast<int4 *>(acc + g * FILL_ACC + c_own * FILL_CS + h_own * 8);
        const int4 ra = own[0], rb = own[1];
        own[0] = make_int4(0, 0, 0, 0);
        own[1] = make_int4(0, 0, 0, 0);
        const int half_total = (ra.x + ra.y + ra.z) + (ra.w + rb.x + rb.y) + (rb.z + rb.w);
        const int above = __shfl_xor_sync(0xffffffffu, half_total, 16);
        int cv[8];
        cv[0] = (int)(int8_t)(at_w & 0xffu) * FILL_ONE + (h_own ? above : 0) + ra.x;
        cv[1] = cv[0] + ra.y; cv[2] = cv[1] + ra.z; cv[3] = cv[2] + ra.w;
        cv[4] = cv[3] + rb.x; cv[5] = cv[4] + rb.y; cv[6] = cv[5] + rb.z; cv[7] = cv[6] + rb.w;
        // round(v * 255) for v in [0, 1] as 12.20 fixed point = (v * 255 + 2^19) >> 20; scaled by 16 it is the TOP BYTE
        // of a 32-bit product, which the byte permutes below pick up directly
        uint32_t bytes[8];
        if (at_x >> 31) {  // winding: min(|cv|, 1)
#pragma unroll
            for (int q = 0; q < 8; q++) bytes[q] = (uint32_t)min(abs(cv[q]), FILL_ONE) * (255u << 4) + (1u << 23);
        } else {           // even-odd: 1 - |1 - mod(cv, 2)|
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int v = FILL_ONE - abs(FILL_ONE - (cv[q] & (2 * FILL_ONE - 1)));
                bytes[q] = (uint32_t)v * (255u << 4) + (1u << 23);
            }
        }
        // this lane's column as two words (rows 8h .. 8h+3, rows 8h+4 .. 8h+7), then the transposes
        uint32_t wa = __byte_perm(__byte_perm(bytes[0], bytes[1], 0x0073), __byte_perm(bytes[2], bytes[3], 0x0073), 0x5410);
        uint32_t wb = __byte_perm(__byte_perm(bytes[4], bytes[5], 0x0073), __byte_perm(bytes[6], bytes[7], 0x0073), 0x5410);
        wa = __byte_perm(wa, __shfl_xor_sync(0xffffffffu, wa, 1), sel1);
        wb = __byte_perm(wb, __shfl_xor_sync(0xffffffffu, wb, 1), sel1);
        wa = __byte_perm(wa, __shfl_xor_sync(0xffffffffu, wa, 2), sel2);
        wb = __byte_perm(wb, __shfl_xor_sync(0xffffffffu, wb, 2), sel2);
        if ((int)at_y >= 0 && at_y < b.mask_capacity) {  // fill.comp:147-150: min() with the clip mask (warp-uniform)
            const uint8_t *clip = b.masks + (size_t)at_y * 256;
            wa = __vminu4(wa, __ldg(reinterpret_cast<const uint32_t *>(clip + off_a)));
            wb = __vminu4(wb, __ldg(reinterpret_cast<const uint32_t *>(clip + off_b)));
        }
        if (CACHE && (code >> 31)) {
            uint8_t *const dst = cache + (size_t)(code & 0x7fffffffu) * 256;
            *reinterpret_cast<uint32_t *>(dst + off_a) = wa;
            *reinterpret_cast<uint32_t *>(dst + off_b) = wb;
        } else {
            uint8_t *const dst = b.masks + (size_t)code * 256;
            *reinterpret_cast<uint32_t *>(dst + off_a) = wa;
            *reinterpret_cast<uint32_t *>(dst + off_b) = wb;
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(FILL_WARPS * 32, FILL_CTAS_PER_SM) k_fill(FillArgs b, LutView p) {
    __shared__ FillShared sh;
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_FILL);
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t first_alpha = b.counters->first_alpha;
    uint32_t n_alpha = b.counters->n_alpha;
    if (n_alpha > b.alpha_capacity) n_alpha = b.alpha_capacity;
    uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
#if FILL_PERMUTE
    // With work for every warp, consecutive groups go to consecutive CTAs (different SMs): the groups of one path --
    // similar fill counts -- are spread over the whole GPU instead of landing on the 8 warps of one CTA (tiger 4096^2:
    // SM-active time 21k .. 42k cycles before). A small frame keeps the CTA-major order, so that most CTAs of the
    // (one-wave) grid exit at once and leave their SMs to the other frames in flight (batch of 512^2 frames: 74k vs
    // 64k frames/s).
    if (n_alpha >= n_warps * FILL_GROUP) warp = wib * gridDim.x + blockIdx.x;
#endif
    int *const acc = &sh.acc[wib][0][0];
    {  // (16-byte stores: the scalar loop was 7 % of the kernel's instructions on tiger 4096^2)
        static_assert((FILL_GROUP * FILL_ACC) % 128 == 0, "whole 16-byte stores per lane");
        uint4 *const acc4 = reinterpret_cast<uint4 *>(acc);
#pragma unroll
        for (int i = 0; i < FILL_GROUP * FILL_ACC / 128; i++) acc4[i * 32 + (int)lane] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncwarp();

    // (Handing the groups out through a global ticket counter instead -- one atomic per group, next ticket prefetched --
    // measured slower: 28.7 us against 24.6 us on tiger 4096^2, profiles/r01_tile_kernel_experiments.md.)
    // Round 2 tried again to hand out the groups beyond every warp's first dynamically, because CTAs do not all start
    // together (tools/timeline.py: the 32 SMs that hold a CTA of the scan over framebuffer tiles get their fill CTAs 6 us late
    // -- an SM keeps its shared-memory carve-out while CTAs are resident -- and those 128 CTAs are the kernel's tail): per
    // warp and one group ahead 32.8 us, per warp on demand 27.6 us, one atomic per CTA for its 8 warps' second groups 26.6 us,
    // against 24.6 us for the fixed assignment below (profiles/r02_tile_kernel.md section 12).
    // Also measured there: the next group's records fetched by cp.async during the current group + L1 prefetches of its z-buffer
    // entries and fills (no change: 24.6 us, and 690 us on the 200 k-path scene -- the loads are not what a round waits for);
    // a grid that leaves out the SMs the scan over framebuffer tiles holds (464 CTAs: some warps take three groups, 30.7 us).
    const uint32_t n_groups = (n_alpha + FILL_GROUP - 1) / FILL_GROUP;
    for (uint32_t grp = warp; grp < n_groups; grp += n_warps) {
        const uint32_t a0 = grp * FILL_GROUP;
        // ---- the group's alpha tile records, one per lane: tile | winding << 31, clip mask slot, first fill,
        // backdrop | fill count << 8
        uint4 at = make_uint4(0x7fffffffu, 0xffffffffu, 0u, 0u);
        if (lane < FILL_GROUP && a0 + lane < n_alpha && first_alpha + a0 + lane < b.mask_capacity)
            at = __ldg(reinterpret_cast<const uint4 *>(&b.alpha_tiles[a0 + lane]));
        uint32_t fbt = 0xffffffffu;
        if (b.cull_fill && lane < FILL_GROUP && a0 + lane < n_alpha) fbt = __ldg(&b.alpha_map[a0 + lane]);
        // a mask under an opaque whole-tile layer of a later path is never read (the list scatter leaves its tile out,
        // sort.comp:62): do not rasterize it. The z-buffer is final: propagate has finished.
        if (fbt != 0xffffffffu && (int)(at.x & 0x7fffffffu) < b.fb[fbt].z) at.x = 0x7fffffffu;
        fill_group<FILL_GROUP, false>(at, first_alpha + a0 + lane, nullptr, acc, b, p, lane);
    }
}

cudaError_t launch_fill(const BatchView &b, const PaintView &p, cudaStream_t s) {
    if (!b.tile_count || !p.lut_tex || b.fused_fill) return cudaSuccess;  // fused: the tile kernel rasterizes the masks
    LutView lv;
    lv.lut_tex = p.lut_tex;
    lv.lut_band = p.lut_band;
    return launch_pdl(k_fill, sm_count() * FILL_CTAS_PER_SM, FILL_WARPS * 32, 0, s, FillArgs(b), lv);
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
// AXIS 1 / 2: the blur runs along x / y only (what the reference's two-pass shadows do, canvas.cpp:91-151): the other
// axis' texel rows / columns and weight are the same for every tap (cv - 0 * k == cv) and are prepared once.
template <int AXIS>
__device__ __forceinline__ float4 filter_blur_loop(const ColorSampler &cs, float cu, float cv, float sox, float soy, float4 p0, float4 p1) {
    const int support = (int)p0.z;
    float gx = p1.x, gy = p1.y;
    const float gz = p1.z;
    float gauss_sum = gx;
    AxisTaps fixed = {};
    if (AXIS == 1) fixed = axis_taps(cv, cs.h, cs.repeat_v);
    if (AXIS == 2) fixed = axis_taps(cu, cs.w, cs.repeat_u);
    auto tap = [&](float u, float v) {
        if (AXIS == 1) return sample_axes(cs.px, cs.w, axis_taps(u, cs.w, cs.repeat_u), fixed);
        if (AXIS == 2) return sample_axes(cs.px, cs.w, fixed, axis_taps(v, cs.h, cs.repeat_v));
        return cs(u, v);
    };
    float4 color = tap(cu, cv);
    color.x *= gx; color.y *= gx; color.z *= gx; color.w *= gx;
    gx *= gy;
    gy *= gz;
    for (int i = 1; i <= support; i += 2) {
        float partial = gx;
        gx *= gy;
        gy *= gz;
        partial += gx;
        const float k = (float)i + gx / partial;
        const float4 a = tap(cu - sox * k, cv - soy * k), bq = tap(cu + sox * k, cv + soy * k);
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

__device__ float4 filter_blur(const ColorSampler &cs, float cu, float cv, float4 p0, float4 p1) {
    const float sox = p0.x / (float)cs.w, soy = p0.y / (float)cs.h;
    if (!cs.nearest && soy == 0.0f) return filter_blur_loop<1>(cs, cu, cv, sox, soy, p0, p1);
    if (!cs.nearest && sox == 0.0f) return filter_blur_loop<2>(cs, cu, cv, sox, soy, p0, p1);
    return filter_blur_loop<0>(cs, cu, cv, sox, soy, p0, p1);
}

// filterColorMatrix, tile.comp:394-404: mat4(p0 .. p3) * texel + p4 (the parameters are the matrix's COLUMNS)
__device__ float4 filter_color_matrix(const ColorSampler &cs, float cu, float cv, const Paint *gp) {
    const float4 s = cs(cu, cv);
    const float4 c0 = __ldg(&gp->fp0), c1 = __ldg(&gp->fp1), c2 = __ldg(&gp->fp2), c3 = __ldg(&gp->fp3), o = __ldg(&gp->fp4);
    return make_float4(c0.x * s.x + c1.x * s.y + c2.x * s.z + c3.x * s.w + o.x, c0.y * s.x + c1.y * s.y + c2.y * s.z + c3.y * s.w + o.y,
                       c0.z * s.x + c1.z * s.y + c2.z * s.z + c3.z * s.w + o.z, c0.w * s.x + c1.w * s.y + c2.w * s.z + c3.w * s.w + o.w);
}

// filterText, tile.comp:136-227, without its gamma correction (the reference binds a 1 x 1 dummy as gamma LUT,
// d3d11/renderer.cpp:262-266; paints that ask for it are refused at upload): coverage from the red channel, optionally
// defringed by a 9-tap / 7-tap-per-channel horizontal kernel, then mix(background, foreground, coverage)
__device__ float4 filter_text(const ColorSampler &cs, float cu, float cv, const Paint *gp) {
    const float4 k = __ldg(&gp->fp0), bg = __ldg(&gp->fp1), fg = __ldg(&gp->fp2);
    float ar, ag, ab;
    if (k.w == 0.0f) {
        ar = ag = ab = cs(cu, cv).x;
    } else {
        const float one = 1.0f / (float)cs.w;
        const bool wide = k.x > 0.0f;
        float t[9];  // taps -4 .. 4
#pragma unroll
        for (int i = 0; i < 9; i++) t[i] = ((i == 0 || i == 8) && !wide) ? 0.0f : cs(cu + (float)(i - 4) * one, cv).x;
        // filterTextConvolve7Tap(alpha0, alpha1, kernel) = dot(alpha0, kernel) + dot(alpha1, kernel.zyx), centred on taps 3, 4, 5
        auto conv = [&](int c) {
            return (t[c - 3] * k.x + t[c - 2] * k.y + t[c - 1] * k.z + t[c] * k.w) + (t[c + 1] * k.z + t[c + 2] * k.y + t[c + 3] * k.x);
        };
        ar = conv(3);
        ag = conv(4);
        ab = conv(5);
    }
    return make_float4(mixf(bg.x, fg.x, ar), mixf(bg.y, fg.y, ag), mixf(bg.z, fg.z, ab), 1.0f);
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
// (not inlined: one copy of the gradient / blur / blend-mode code per kernel instead of one per pixel and call site --
// inlining it made the textured instantiation of the tile kernel 1 MB of SASS, an instruction-cache disaster)
template <bool SOLID>
__device__ __noinline__ float4 shade(const Paint &pc, const Paint *gp, const ColorSampler &cs, float fragx, float fragy,
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
            else if (filter == 0x2) c0 = filter_text(cs, cu, cv, gp);
            else if (filter == 0x4) c0 = filter_color_matrix(cs, cu, cv, gp);
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

#ifndef CT_WARPS_N
#define CT_WARPS_N 4
#endif
#ifndef CT_MIN_CTAS
#define CT_MIN_CTAS (28 / CT_WARPS_N)  // 7 CTAs per SM = 72 registers per thread (6: 78 registers measured 2 us slower; 64 registers spill)
#endif
#ifndef CT_MIN_CTAS_TEX
#define CT_MIN_CTAS_TEX 4  // gradients / images / blend modes: up to 168 registers
#endif
constexpr int CT_WARPS = CT_WARPS_N;   // warps per CTA
constexpr int CT_THREADS = CT_WARPS * 32;
constexpr int CT_TILES = CT_WARPS * 4; // consecutive framebuffer tiles of one group (8 threads order one tile's list)
#if CT_WARPS_N == 4
static_assert(CT_TILES == GROUP_TILES, "the scan lays the headers out for groups of CT_TILES tiles");
#endif
constexpr int CT_PRIMS = 256;          // list entries staged per group (more: the CTA takes the group in several rounds)
constexpr uint32_t KEY_MASK = 0x00ffffffu;
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

// base colour of a resolved list entry (TilePrim): four halfs -> floats, exact
__device__ __forceinline__ float4 prim_color(const uint4 u) {
    const float2 rg = __half22float2(*reinterpret_cast<const __half2 *>(&u.z));
    const float2 ba = __half22float2(*reinterpret_cast<const __half2 *>(&u.w));
    return make_float4(rg.x, rg.y, ba.x, ba.y);
}

// The lane's 8 pixels, channel-major in pairs: pair j holds pixels 2j and 2j+1 (pixel q = column 4k + (q & 3) of row
// 2s + (q >> 2)), so every blend is a packed f32x2 operation with a per-layer constant or a per-pixel alpha.
struct PixelBlock {
    float2 x[4], y[4], z[4], w[4];
    __device__ __forceinline__ void set_all(float4 c) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x[j] = splat(c.x);
            y[j] = splat(c.y);
            z[j] = splat(c.z);
            w[j] = splat(c.w);
        }
    }
    __device__ __forceinline__ void set(int q, float4 c) {
        if (q & 1) { x[q >> 1].y = c.x; y[q >> 1].y = c.y; z[q >> 1].y = c.z; w[q >> 1].y = c.w; }
        else { x[q >> 1].x = c.x; y[q >> 1].x = c.y; z[q >> 1].x = c.z; w[q >> 1].x = c.w; }
    }
    // dest = dest * (1 - src.a) + src for one premultiplied colour over all 8 pixels
    __device__ __forceinline__ void over_all(float4 src) {
        const float2 ia = splat(1.0f - src.w), sx = splat(src.x), sy = splat(src.y), sz = splat(src.z), sw = splat(src.w);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            x[j] = fma2(x[j], ia, sx);
            y[j] = fma2(y[j], ia, sy);
            z[j] = fma2(z[j], ia, sz);
            w[j] = fma2(w[j], ia, sw);
        }
    }
    // pair j: premultiplied sources (sx, sy, sz, sw) of the two pixels
    __device__ __forceinline__ void over_pair(int j, float2 sx, float2 sy, float2 sz, float2 sw) {
        const float2 ia = fma2(sw, splat(-1.0f), splat(1.0f));  // 1 - a, one rounding like the scalar subtraction
        x[j] = fma2(x[j], ia, sx);
        y[j] = fma2(y[j], ia, sy);
        z[j] = fma2(z[j], ia, sz);
        w[j] = fma2(w[j], ia, sw);
    }
    // RGBA8 of pixels 4r .. 4r+3 (one 16-byte row piece)
    template <bool SAT>
    __device__ __forceinline__ uint4 pack_row(int r) const {
        uint32_t wd[4];
#pragma unroll
        for (int jj = 0; jj < 2; jj++) {
            const int j = r * 2 + jj;
            float2 vx = x[j], vy = y[j], vz = z[j], vw = w[j];
            if (SAT) {
                vx = make_float2(__saturatef(vx.x), __saturatef(vx.y));
                vy = make_float2(__saturatef(vy.x), __saturatef(vy.y));
                vz = make_float2(__saturatef(vz.x), __saturatef(vz.y));
                vw = make_float2(__saturatef(vw.x), __saturatef(vw.y));
            }
            const float2 k = splat(255.0f), m = splat(12582912.0f);  // round to nearest even into the mantissa
            const float2 bx = fma2(vx, k, m), by = fma2(vy, k, m), bz = fma2(vz, k, m), bw = fma2(vw, k, m);
            wd[jj * 2 + 0] = __byte_perm(__byte_perm(__float_as_uint(bx.x), __float_as_uint(by.x), 0x0040),
                                         __byte_perm(__float_as_uint(bz.x), __float_as_uint(bw.x), 0x0040), 0x5410);
            wd[jj * 2 + 1] = __byte_perm(__byte_perm(__float_as_uint(bx.y), __float_as_uint(by.y), 0x0040),
                                         __byte_perm(__float_as_uint(bz.y), __float_as_uint(bw.y), 0x0040), 0x5410);
        }
        return make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
};

// coverage of pixels 2j, 2j+1 from the lane's 8 mask bytes: float(byte) / 255 like the RGBA8-unorm fetch of
// tile.comp:594-598 (the byte lands in the mantissa of 2^23, the subtraction is exact)
__device__ __forceinline__ float2 mask_pair(uint2 mask8, int j, bool even_odd) {
    const uint32_t m = j < 2 ? mask8.x : mask8.y;
    const int i0 = (j & 1) * 2;
    float2 c = make_float2(__uint_as_float(__byte_perm(m, 0x4b000000u, 0x7540 | i0)),
                           __uint_as_float(__byte_perm(m, 0x4b000000u, 0x7540 | (i0 + 1))));
    c = mul2(add2(c, splat(-8388608.0f)), splat(1.0f / 255.0f));
    // sampleMask's even-odd fold, tile.comp:601-602: 1 - |1 - mod(c, 2)| with c in [0, 1]
    if (even_odd) c = fma2(fma2(c, splat(-1.0f), splat(1.0f)), splat(-1.0f), splat(1.0f));
    return c;
}

#ifndef FUSE_GROUP_N
#define FUSE_GROUP_N 2
#endif
#ifndef FUSE_MASKS_N
#define FUSE_MASKS_N 32
#endif
constexpr int FUSE_GROUP = FUSE_GROUP_N;  // masks a warp of the tile kernel rasterizes together
constexpr int FUSE_MASKS = FUSE_MASKS_N;  // masks of a tile group kept in shared memory (the others go through b.masks)

// What the tile kernel needs on top of its own staging buffers to rasterize the masks of its tiles itself
// (BatchView::fused_fill): accumulators, the finished masks, the alpha tile records fetched while the lists are ordered.
template <bool FUSED>
struct __align__(16) FusedShared {};
template <>
struct __align__(16) FusedShared<true> {
    int acc[CT_WARPS][FUSE_GROUP][FILL_ACC];
    uint2 cache[FUSE_MASKS][32];
    uint4 at[FUSE_MASKS];
    uint32_t alpha[CT_PRIMS];  // batch-local alpha tile of every mask the group needs
    uint32_t n_masks, next_mask;
};

template <bool SOLID, bool FUSED>
struct __align__(16) CompositeShared {
    uint4 fb[CT_TILES];                // list begin, slots (count before z-cull), z, entries
    uint4 raw[CT_PRIMS];               // the lists as the scatter left them
    uint4 sorted[CT_PRIMS];            // in paint order, resolved: key | LayerFlags << 24, mask slot, colour (4 halfs)
    float4 start_color[CT_TILES];      // the tile's colour after its leading whole-tile layers
    uint32_t start_layer[CT_TILES];    // first layer that needs per-pixel work
    uint32_t packed_color[CT_TILES];
    uint32_t txy[CT_TILES];            // tile x | tile y << 16
    uint16_t paint[SOLID ? 8 : CT_PRIMS]; // paint of every layer (textured layers look their constants up at blend time)
    uint16_t work[CT_TILES * 4];       // per-pixel work items: tile | pixel pairs (bit j = pair j of every lane) << 8
    uint32_t n_work, next, flat_mask;
    FusedShared<FUSED> fz;
};

struct TileGeom {
    int gx0, gy0;       // the lane's first pixel
    bool interior;      // the whole tile is inside the target
    uint8_t *px0;       // address of the lane's first pixel
    float fragx, fragy; // gl_FragCoord of the lane's first pixel on the full canvas
};

// A general-format list entry (key, mask slot, paint | ctrl << 16 | backdrop << 24) resolved against the paint table:
// key | LayerFlags << 24, mask slot, base colour as halfs. Frames whose paints are all plain colours get this from the list
// scatter already (BatchView::solid_prims).
template <bool SOLID, class B>
__device__ __forceinline__ uint4 resolve_prim(const uint4 q, const B &b, const PaintView &p, uint32_t &paint) {
    if (SOLID || b.solid_prims) {
        paint = 0;
        return q;
    }
    paint = q.z & 0xffffu;
    uint32_t fl = layer_flags((int)q.y, (q.z >> 16) & 0xffu, (int)q.z >> 24, b.mask_capacity);
    float4 base = make_float4(0.f, 0.f, 0.f, 0.f);
    if (paint < p.n_paints) {
        base = __ldg(&p.paints[paint].base);
        const int ctrl = __ldg(&p.paints[paint].ctrl);
        if (ctrl != 0) fl |= LF_TEXTURED;
        if (((ctrl >> 8) & 0x3) != 0 && ((ctrl >> 4) & 0xf) == 0x3) fl |= LF_HEAVY;
    }
    const __half2 rg = __floats2half2_rn(base.x, base.y), ba = __floats2half2_rn(base.z, base.w);
    return make_uint4((q.x & KEY_MASK) | (fl << 24), q.y, *reinterpret_cast<const uint32_t *>(&rg), *reinterpret_cast<const uint32_t *>(&ba));
}

// the lane's 8 coverage bytes of a layer (all ones for a layer without a mask). FUSED: the masks this kernel rasterized
// itself are in shared memory (slot | 1 << 31) or were written to b.masks by this CTA (no read-only path for those)
template <bool FUSED, class B>
__device__ __forceinline__ uint2 load_mask(const uint4 u, const B &b, const uint2 (*cache)[32], unsigned lane) {
    if (!(u.x & ((uint32_t)LF_MASKED << 24)) || (u.x & ((uint32_t)LF_SKIP << 24))) return make_uint2(0xffffffffu, 0xffffffffu);
    if (FUSED) {
        if (u.y >> 31) return cache[u.y & 0x7fffffffu][lane];
        return __ldcg(reinterpret_cast<const uint2 *>(b.masks + (size_t)u.y * 256) + lane);
    }
    return __ldg(reinterpret_cast<const uint2 *>(b.masks + (size_t)u.y * 256) + lane);
}

// One layer over the lane's 8 pixels (tile.comp:765-842 for one list entry).
template <bool SOLID, class B>
__device__ __forceinline__ void blend_layer(PixelBlock &px, const uint4 u, const uint2 mask8, uint32_t paint, const B &b,
                                            const PaintView &p, const ColorSampler &cs, const TileGeom &g,
                                            const TargetView &tg, unsigned lane, uint32_t pairs = 0xfu) {
    const uint32_t fl = u.x >> 24;
    if (fl & LF_SKIP) return;
    const bool masked = (fl & LF_MASKED) != 0, textured = !SOLID && (fl & LF_TEXTURED) != 0;
    const float4 base = prim_color(u);
    if (!masked && !textured) {  // one premultiplied colour over the whole tile (calculateColor with maskAlpha == 1)
        px.over_all(make_float4(base.x * base.w, base.y * base.w, base.z * base.w, base.w));
        return;
    }
    const bool even_odd = (fl & LF_EVEN_ODD) != 0;
    if (!textured) {
        const float2 bx = splat(base.x), by = splat(base.y), bz = splat(base.z), bw = splat(base.w);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float2 a = mul2(mask_pair(mask8, j, even_odd), bw);  // color.a *= maskAlpha, tile.comp:672
            px.over_pair(j, mul2(bx, a), mul2(by, a), mul2(bz, a), a);
        }
        return;
    }
    if (!SOLID) {
        Paint pc;
        pc.base = base;
        pc.m0 = __ldg(&p.paints[paint].m0);
        pc.m1 = __ldg(&p.paints[paint].m1);
        pc.fp0 = __ldg(&p.paints[paint].fp0);
        pc.fp1 = __ldg(&p.paints[paint].fp1);
        pc.ctrl = __ldg(&p.paints[paint].ctrl);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (!((pairs >> j) & 1u)) continue;  // (a tile split over several warps: the other pairs are theirs)
            const float2 cov = mask_pair(mask8, j, even_odd);
            const float fy = g.fragy + (float)(j >> 1), fx = g.fragx + (float)((j & 1) * 2);
            const float4 s0 = shade<false>(pc, &p.paints[paint], cs, fx, fy, cov.x, (float)tg.width, (float)tg.height);
            const float4 s1 = shade<false>(pc, &p.paints[paint], cs, fx + 1.0f, fy, cov.y, (float)tg.width, (float)tg.height);
            px.over_pair(j, make_float2(s0.x, s1.x), make_float2(s0.y, s1.y), make_float2(s0.z, s1.z),
                         make_float2(s0.w, s1.w));
        }
    }
}

template <bool SOLID>
__device__ __forceinline__ void store_block(const PixelBlock &px, const TileGeom &g, const TargetView &tg,
                                            uint32_t pairs = 0xfu) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint4 v = px.pack_row<!SOLID>(r);
        uint8_t *dst = g.px0 + (size_t)r * tg.pitch;
        if (!SOLID && pairs != 0xfu) {  // this warp owns some of the lane's pixel pairs only
            const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (((pairs >> (r * 2 + (i >> 1))) & 1u) && g.gx0 + i < tg.width && g.gy0 + r < tg.height)
                    reinterpret_cast<uint32_t *>(dst)[i] = wd[i];
        } else if (g.interior) {
            *reinterpret_cast<uint4 *>(dst) = v;
        } else if (g.gy0 + r < tg.height) {
            const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (g.gx0 + i < tg.width) reinterpret_cast<uint32_t *>(dst)[i] = wd[i];
        }
    }
}

__device__ __forceinline__ void load_block(PixelBlock &px, const TileGeom &g, const TargetView &tg) {
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint8_t *src = g.px0 + (size_t)r * tg.pitch;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const bool ok = g.interior || (g.gx0 + i < tg.width && g.gy0 + r < tg.height);
            px.set(r * 4 + i, ok ? unpack_rgba8(reinterpret_cast<const uint32_t *>(src)[i]) : make_float4(0.f, 0.f, 0.f, 0.f));
        }
    }
}

__device__ __forceinline__ TileGeom tile_geom(uint32_t xy, unsigned lane, bool vec_ok, float org_x, float org_y, const TargetView &tg) {
    TileGeom g;
    const int tile_x = (int)(xy & 0xffffu), tile_y = (int)(xy >> 16);
    g.gx0 = tile_x * TILE + (int)(lane & 3u) * 4;
    g.gy0 = tile_y * TILE + (int)(lane >> 2) * 2;
    g.interior = vec_ok && (tile_x + 1) * TILE <= tg.width && (tile_y + 1) * TILE <= tg.height;
    g.px0 = tg.pixels + (size_t)g.gy0 * tg.pitch + (size_t)g.gx0 * 4;
    g.fragx = (float)g.gx0 + org_x;
    g.fragy = (float)g.gy0 + org_y;
    return g;
}


__device__ __forceinline__ uint4 clamp_header(uint4 f, uint32_t prim_capacity) {
    if (f.x > prim_capacity) f.x = prim_capacity;
    if (f.x + f.y > prim_capacity) f.y = prim_capacity - f.x;
    f.w = min(f.w, f.y);  // entries the scatter wrote (it leaves out what the z-buffer culls)
    return f;
}

// One CTA renders CT_TILES consecutive framebuffer tiles. The dependent loads of a tile (list header -> list -> masks) are
// issued for all of the CTA's tiles at once, so their latency is paid once per CTA; the lists are ordered and classified by
// all threads (8 per tile); then the whole CTA stores the one-colour tiles and the warps pull the remaining tiles from a
// shared counter, which balances deep lists against shallow ones.
// Measured alternatives that lost on tiger 4096^2 (profiles/r01_tile_kernel_experiments.md, profiles/r02_tile_kernel.md):
// persistent grids (device-wide ticket; round 2: headers and lists prefetched two groups ahead by bulk copies into a
// shared-memory ring -- the loads no longer stall anybody, the kernel is still slower, 43 us); 32 / 64 tiles per CTA with
// headers and lists staged once by bulk copies (39 / 48 us); one warp per 2 / 4 / 8 tiles with no block-wide barrier at all
// (41 - 50 us); a separate sort kernel + warp-per-tile blend kernel; streaming / evict-first / TMA tensor stores for the
// framebuffer (no faster than plain 16-byte stores, tools/ubench/store_floor.cu).
//
// FUSED (BatchView::fused_fill, PFCU_OPT_FUSED_FILL): the fill stage runs INSIDE this kernel. The layers that need a mask of
// this batch are collected while the lists are ordered (their alpha tile records fetched, their fills prefetched); before
// the per-pixel tiles are blended the CTA's warps rasterize those masks, FUSE_GROUP at a time, with fill's own code
// (fill_group) into shared memory. Only masks that a list references are ever rasterized, none is written to or read from
// HBM, and k_fill is not launched for the batch. Masks of other batches (clip masks) are read from b.masks as before.
// Byte-identical frames, but measured SLOWER than the separate fill kernel on tiger 4096^2 (64.2 us against 24.6 us of fill
// beside the list building + 29.1 us: 23.1 M warp instructions, SM-active 64 K .. 115 K of 118 K elapsed cycles -- the masked
// tiles of a group serialise behind their CTA's four warps, and a quarter of the warp time waits at the barrier between
// rasterizing and blending; profiles/r02_tile_kernel.md section 7), so it is off by default.
template <bool SOLID, bool FUSED>
__global__ void __launch_bounds__(CT_THREADS, SOLID ? CT_MIN_CTAS : CT_MIN_CTAS_TEX) k_composite(CompositeArgs b, PaintView p, TargetView tg, int clear,
                                                                      float4 clear_color, int origin, uint32_t tiles_per_cta,
                                                                      uint32_t sub_tw, uint32_t sub_n, int ordered) {
    __shared__ CompositeShared<SOLID, FUSED> sh;
    PFCU_KERNEL_BEGIN(b, PFCU_STAGE_COMPOSITE);
    const unsigned tid = threadIdx.x, lane = tid & 31;
    const uint2(*mask_cache)[32] = nullptr;
    uint32_t first_alpha = 0, n_alpha = 0;
    bool acc_zeroed = false;
    if constexpr (FUSED) {
        mask_cache = sh.fz.cache;
        first_alpha = __ldg(&b.counters->first_alpha);
        n_alpha = min(__ldg(&b.counters->n_alpha), b.alpha_capacity);
    }
    // the CTA's tiles: tiles_per_cta consecutive tiles of the sub_tw-wide rectangle of the tile grid that the target
    // covers (the whole grid for the destination; a render-target page is smaller than the scene's grid)
    // `ordered` (the destination, GROUP_TILES tiles per CTA, groups sorted by cost: BatchView::fb_sorted): this CTA renders
    // group group_of[blockIdx.x], whose headers the scan left at blockIdx.x * GROUP_TILES -- both loads go out together.
    uint4 f_direct = make_uint4(0u, 0u, 0u, 0u);
    if (ordered && tid < CT_TILES) f_direct = __ldg(reinterpret_cast<const uint4 *>(&b.fb_sorted[blockIdx.x * GROUP_TILES + tid]));
    const uint32_t map0 = (ordered ? __ldg(&b.group_of[blockIdx.x]) : blockIdx.x) * tiles_per_cta;
    const uint32_t n_tiles = min(tiles_per_cta, sub_n - map0);

    // ---- stage 1: list headers
    if (tid < CT_TILES) {
        uint4 f = make_uint4(0u, 0u, 0u, 0u);
        uint32_t xy = 0;
        if (tid < n_tiles) {
            const uint32_t at = map0 + tid, ty = at / sub_tw, tx = at - ty * sub_tw;
            if (!ordered) f_direct = __ldg(reinterpret_cast<const uint4 *>(fb_header(b, ty * (uint32_t)b.fb_tw + tx)));
            f = clamp_header(f_direct, b.prim_capacity);
            xy = tx | (ty << 16);
        }
        sh.fb[tid] = f;
        sh.txy[tid] = xy;
    }
    __syncthreads();

    ColorSampler cs;
    cs.px = p.color_px;
    cs.w = p.color_w;
    cs.h = p.color_h;
    cs.repeat_u = (p.sampling_flags & 1u) != 0;
    cs.repeat_v = (p.sampling_flags & 2u) != 0;
    cs.nearest = (p.sampling_flags & 0xcu) != 0;
    // 16-byte stores need 16-byte rows (a render-target page of odd width has none: 4-byte stores there)
    const bool vec_ok = ((reinterpret_cast<size_t>(tg.pixels) | tg.pitch) & 15) == 0;
    // gl_FragCoord of the full canvas (only textured paints look at it); render-target pages have no origin
    const float org_x = (float)(origin ? b.fb_tx0 * TILE : 0) + 0.5f, org_y = (float)(origin ? b.fb_ty0 * TILE : 0) + 0.5f;

    uint32_t tb = 0;
    while (tb < n_tiles) {
        // ---- the tiles of this round: [tb, te), as many as the staging buffers hold (normally all of them)
        const uint32_t range0 = sh.fb[tb].x;
        uint32_t te = n_tiles;
        if (sh.fb[n_tiles - 1].x + sh.fb[n_tiles - 1].y - range0 > CT_PRIMS) {
            te = tb + 1;
            while (te < n_tiles && sh.fb[te].x + sh.fb[te].y - range0 <= CT_PRIMS) te++;
        }
        const uint32_t range_n = sh.fb[te - 1].x + sh.fb[te - 1].y - range0;
        if (tid == 0) {
            sh.n_work = 0;
            sh.next = 0;
            sh.flat_mask = 0;
            if constexpr (FUSED) sh.fz.n_masks = sh.fz.next_mask = 0;
        }
        if (range_n > CT_PRIMS) {
            // ---- a single tile whose list does not fit (te == tb + 1): one warp walks it by repeated selection of the
            // next key from global memory
            __syncthreads();
            if (tid < 32) {
                const uint32_t t = tb;
                const uint4 hdr = sh.fb[t];
                const TileGeom g = tile_geom(sh.txy[t], lane, vec_ok, org_x, org_y, tg);
                const uint4 *list = reinterpret_cast<const uint4 *>(b.prims) + hdr.x;
                auto next_layer = [&](uint32_t &last_key, bool &first, uint4 &u, uint32_t &paint) {
                    uint32_t best = 0xffffffffu, best_i = 0;
                    for (uint32_t i = lane; i < hdr.w; i += 32) {
                        const uint32_t key = __ldg(&list[i].x) & KEY_MASK;
                        if ((first || key > last_key) && key < best) {
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
                    last_key = best;
                    first = false;
                    u = resolve_prim<SOLID>(__ldg(&list[best_i]), b, p, paint);
                    return true;
                };
                if constexpr (FUSED) {
                    // the list's own masks first, through b.masks (where k_fill would have put them)
                    int *const acc = &sh.fz.acc[0][0][0];
                    if (!acc_zeroed) {
                        for (int i = (int)lane; i < FUSE_GROUP * FILL_ACC; i += 32) acc[i] = 0;
                        acc_zeroed = true;
                        __syncwarp();
                    }
                    for (uint32_t i0 = 0; i0 < hdr.w; i0 += FUSE_GROUP) {
                        uint4 at = make_uint4(0x7fffffffu, 0xffffffffu, 0u, 0u);
                        uint32_t code = 0;
                        if (lane < FUSE_GROUP && i0 + lane < hdr.w) {
                            uint32_t paint;
                            const uint4 u = resolve_prim<SOLID>(__ldg(&list[i0 + lane]), b, p, paint);
                            const uint32_t a = u.y - first_alpha;
                            if ((u.x & ((uint32_t)LF_MASKED << 24)) && u.y >= first_alpha && a < n_alpha) {
                                at = __ldg(reinterpret_cast<const uint4 *>(&b.alpha_tiles[a]));
                                code = u.y;
                            }
                        }
                        fill_group<FUSE_GROUP, false>(at, code, nullptr, acc, b, p, lane);
                    }
                    __threadfence_block();
                    __syncwarp();
                }
                {
                    PixelBlock px;
                    if (clear) px.set_all(clear_color);
                    else load_block(px, g, tg);
                    uint32_t last_key = 0, paint;
                    bool first = true;
                    uint4 u;
                    while (next_layer(last_key, first, u, paint))
                        blend_layer<SOLID>(px, u, load_mask<FUSED>(u, b, mask_cache, lane), paint, b, p, cs, g, tg, lane);
                    if (clear || hdr.w) store_block<SOLID>(px, g, tg);
                }
            }
            tb = te;
            __syncthreads();
            continue;
        }

        // ---- stage 2: the lists of the round's tiles are one contiguous range
        {
            const uint4 *src = reinterpret_cast<const uint4 *>(b.prims) + range0;
            for (uint32_t i = tid; i < range_n; i += CT_THREADS) sh.raw[i] = __ldg(src + i);
        }
        __syncthreads();

        // ---- stage 3: 8 threads per tile put its list in paint order (keys are unique: rank = number of smaller keys),
        // resolve the layers and start the mask loads; then one of them blends the leading whole-tile layers
        {
            const uint32_t t = tb + (tid >> 3), sub = tid & 7u;
            uint32_t n = 0, off = 0;
            if (t < te) {
                const uint4 hdr = sh.fb[t];
                n = hdr.w;
                off = hdr.x - range0;
                for (uint32_t e = sub; e < n; e += 8) {
                    const uint4 q = sh.raw[off + e];
                    const uint32_t key = q.x & KEY_MASK;
                    uint32_t rank = 0;
                    for (uint32_t j = 0; j < n; j++) rank += (sh.raw[off + j].x & KEY_MASK) < key ? 1u : 0u;
                    uint32_t paint;
                    uint4 u = resolve_prim<SOLID>(q, b, p, paint);
                    if ((u.x & ((uint32_t)LF_MASKED << 24))) {
                        bool own = false;
                        if constexpr (FUSED) {
                            // a mask of this batch: this CTA rasterizes it (stage 3b). Fetch its record, start on its fills
                            const uint32_t a = u.y - first_alpha;
                            own = u.y >= first_alpha && a < n_alpha;
                            if (own) {
                                const uint32_t m = atomicAdd(&sh.fz.n_masks, 1u);
                                sh.fz.alpha[m] = a;
                                if (m < FUSE_MASKS) {
                                    const uint4 at = __ldg(reinterpret_cast<const uint4 *>(&b.alpha_tiles[a]));
                                    sh.fz.at[m] = at;
                                    if (at.z < b.fill_capacity) prefetch_l1(&b.fills[at.z]);
                                    u.y = 0x80000000u | m;
                                }
                            }
                        }
                        if (!own) {
                            prefetch_l1(b.masks + (size_t)u.y * 256);
                            prefetch_l1(b.masks + (size_t)u.y * 256 + 128);
                        }
                    }
                    sh.sorted[off + rank] = u;
                    if (!SOLID) sh.paint[off + rank] = (uint16_t)paint;
                }
            }
            __syncwarp();
            if (t < te && sub == 0) {
                bool is_flat = false, is_work = false;
                float4 dest = clear_color;
                uint32_t i = 0;
                if (n == 0) {
                    is_flat = clear != 0;  // LOAD_ACTION_LOAD leaves an empty tile alone (tile.comp:743-744)
                } else if (!clear) {
                    is_work = true;
                } else {
                    for (; i < n; i++) {
                        const uint4 u = sh.sorted[off + i];
                        const uint32_t fl = u.x >> 24;
                        if (fl & LF_SKIP) continue;
                        if (fl & (LF_MASKED | LF_TEXTURED)) break;
                        float4 src = prim_color(u);
                        src.x *= src.w;
                        src.y *= src.w;
                        src.z *= src.w;
                        blend_over(dest, src);
                    }
                    is_flat = i == n;
                    is_work = !is_flat;
                }
                if (is_flat) {
                    sh.packed_color[t] = pack_rgba8(dest);
                    atomicOr(&sh.flat_mask, 1u << (t - tb));
                } else if (is_work) {
                    sh.start_color[t] = dest;
                    sh.start_layer[t] = i;
                    bool heavy = false;
                    if (!SOLID)
                        for (uint32_t j = clear ? i : 0u; j < n; j++) heavy = heavy || ((sh.sorted[off + j].x >> 24) & LF_HEAVY) != 0;
                    if (heavy) {  // four work items, one pixel pair of every lane each
                        const uint32_t at = atomicAdd(&sh.n_work, 4u);
                        for (uint32_t j = 0; j < 4; j++) sh.work[at + j] = (uint16_t)(t | ((1u << j) << 8));
                    } else {
                        sh.work[atomicAdd(&sh.n_work, 1u)] = (uint16_t)(t | (0xfu << 8));
                    }
                }
            }
        }
        __syncthreads();

        // ---- stage 4a: one-colour tiles, stored by the whole CTA with 16-byte stores. A thread keeps one tile and one
        // 16-byte column of it and walks down the rows; consecutive threads write consecutive 16-byte pieces of a pixel
        // row across the tiles (contiguous when the tiles share a tile row)
        {
            constexpr uint32_t ROW_STEP = CT_THREADS / (4 * CT_TILES);
            const uint32_t tt = (tid >> 2) % CT_TILES, quarter = tid & 3u;
            if ((sh.flat_mask >> tt) & 1u) {
                const uint32_t t = tb + tt, xy = sh.txy[t], c = sh.packed_color[t];
                const int tile_x = (int)(xy & 0xffffu), tile_y = (int)(xy >> 16);
                const int gx = tile_x * TILE + (int)quarter * 4;
                int gy = tile_y * TILE + (int)(tid / (4 * CT_TILES));
                uint8_t *dst = tg.pixels + (size_t)gy * tg.pitch + (size_t)gx * 4;
                const uint4 v = make_uint4(c, c, c, c);
                if (vec_ok && (tile_x + 1) * TILE <= tg.width && (tile_y + 1) * TILE <= tg.height) {
#pragma unroll
                    for (uint32_t r = 0; r < TILE / ROW_STEP; r++, dst += ROW_STEP * tg.pitch) *reinterpret_cast<uint4 *>(dst) = v;
                } else {
                    for (uint32_t r = 0; r < TILE / ROW_STEP; r++, dst += ROW_STEP * tg.pitch, gy += ROW_STEP) {
                        if (gy >= tg.height) break;
                        for (int i = 0; i < 4; i++)
                            if (gx + i < tg.width) reinterpret_cast<uint32_t *>(dst)[i] = c;
                    }
                }
            }
        }

        // ---- stage 3b (FUSED): the masks of this batch that the round's lists reference, FUSE_GROUP per warp and turn
        if constexpr (FUSED) {
            const uint32_t n_masks = sh.fz.n_masks;
            if (n_masks) {
                int *const acc = &sh.fz.acc[tid >> 5][0][0];
                if (!acc_zeroed) {
                    for (int i = (int)lane; i < FUSE_GROUP * FILL_ACC; i += 32) acc[i] = 0;
                    acc_zeroed = true;
                    __syncwarp();
                }
                while (true) {
                    uint32_t m0 = 0;
                    if (lane == 0) m0 = atomicAdd(&sh.fz.next_mask, (uint32_t)FUSE_GROUP);
                    m0 = __shfl_sync(0xffffffffu, m0, 0);
                    if (m0 >= n_masks) break;
                    uint4 at = make_uint4(0x7fffffffu, 0xffffffffu, 0u, 0u);
                    uint32_t code = 0;
                    const uint32_t m = m0 + lane;
                    if (lane < FUSE_GROUP && m < n_masks) {
                        const uint32_t a = sh.fz.alpha[m];
                        if (m < FUSE_MASKS) {
                            at = sh.fz.at[m];
                            code = 0x80000000u | m;
                        } else if (first_alpha + a < b.mask_capacity) {
                            at = __ldg(reinterpret_cast<const uint4 *>(&b.alpha_tiles[a]));
                            code = first_alpha + a;
                        }
                    }
                    fill_group<FUSE_GROUP, true>(at, code, reinterpret_cast<uint8_t *>(sh.fz.cache), acc, b, p, lane);
                }
                __syncthreads();  // (block-uniform: n_masks is)
            }
        }

        // ---- stage 4b: the other tiles, one warp per tile; the mask of layer i + 1 is fetched before layer i is blended
        {
            const uint32_t n_work = sh.n_work;
            while (true) {
                uint32_t wi = 0;
                if (lane == 0) wi = atomicAdd(&sh.next, 1u);
                wi = __shfl_sync(0xffffffffu, wi, 0);
                if (wi >= n_work) break;
                const uint32_t item = sh.work[wi], t = item & 0xffu;
                const uint4 hdr = sh.fb[t];
                const uint32_t n = hdr.w, off = hdr.x - range0;
                const uint32_t i0 = clear ? sh.start_layer[t] : 0u;  // (a work tile has a layer at i0)
                const TileGeom g = tile_geom(sh.txy[t], lane, vec_ok, org_x, org_y, tg);
                const uint32_t pairs = SOLID ? 0xfu : item >> 8;
                uint32_t i = i0;
                uint4 u = sh.sorted[off + i];
                uint2 mask8 = load_mask<FUSED>(u, b, mask_cache, lane);
                PixelBlock px;
                if (clear) px.set_all(sh.start_color[t]);
                else load_block(px, g, tg);
                while (true) {
                    uint4 un = u;
                    uint2 mn = mask8;
                    if (i + 1 < n) {
                        un = sh.sorted[off + i + 1];
                        mn = load_mask<FUSED>(un, b, mask_cache, lane);
                    }
                    blend_layer<SOLID>(px, u, mask8, SOLID ? 0u : (uint32_t)sh.paint[off + i], b, p, cs, g, tg, lane, pairs);
                    if (++i >= n) break;
                    u = un;
                    mask8 = mn;
                }
                store_block<SOLID>(px, g, tg, pairs);
            }
        }
        tb = te;
        if (tb < n_tiles) __syncthreads();  // the next round reuses the staging buffers
    }
}

cudaError_t launch_composite(const BatchView &b, const PaintView &p, const TargetView &t, int clear,
                             const float clear_color[4], int origin, int heavy_paints, cudaStream_t s, int *n_launched) {
    if (n_launched) *n_launched = 0;
    if (b.fb_tw <= 0 || b.fb_th <= 0) return cudaSuccess;
    // tiles of the scene's grid that the target covers (its top-left corner is the grid's: pages have no origin)
    const uint32_t sub_tw = (uint32_t)min(b.fb_tw, (t.width + TILE - 1) / TILE);
    const uint32_t sub_th = (uint32_t)min(b.fb_th, (t.height + TILE - 1) / TILE);
    const uint32_t n_fb = sub_tw * sub_th;
    if (!n_fb) return cudaSuccess;
    if (n_launched) *n_launched = 1;
    const float4 cc = make_float4(clear_color[0], clear_color[1], clear_color[2], clear_color[3]);
    // the plain-colour instantiation skips the saturation before the RGBA8 conversion: src-over of premultiplied colours
    // in [0, 1] stays in [0, 1]
    // CTA j renders the j-th most expensive group: only when the pass walks the whole grid in groups of GROUP_TILES
    // (the destination); other passes find their headers through slot_of
    const int whole = b.fb_sorted && sub_tw == (uint32_t)b.fb_tw && sub_th == (uint32_t)b.fb_th;
    int ordered = whole;
    bool unit = p.all_solid && p.unit_range && b.solid_prims;
    for (int i = 0; i < 4; i++) unit = unit && clear_color[i] >= 0.0f && clear_color[i] <= 1.0f;
    if (unit) {
        uint32_t tpc = CT_TILES;
        // A frame of 4 K .. 8 K tiles (1024^2 .. 1448^2) is at most half a wave of 16-tile CTAs: groups of 8 (tiger.svg @ 1024^2:
        // frame alone 77.8 -> 69.6 us; 4: 67.6 us). Smaller frames keep 16: they come in batches, many contexts side by side, and
        // fewer CTAs per frame serve those better (4096 x 512^2: 99.6 k frames/s against 98.3 k / 95.5 k with groups of 8 / 4).
        if (n_fb >= 4096u && n_fb <= 8192u) {
            tpc = 8;
            ordered = 0;
        }
        if (b.fused_fill)
            return launch_pdl(k_composite<true, true>, (n_fb + tpc - 1) / tpc, CT_THREADS, 0, s, CompositeArgs(b), p, t, clear, cc, origin, tpc, sub_tw, n_fb, ordered);
        return launch_pdl(k_composite<true, false>, (n_fb + tpc - 1) / tpc, CT_THREADS, 0, s, CompositeArgs(b), p, t, clear, cc, origin, tpc, sub_tw, n_fb, ordered);
    }
    // Textured passes on small targets (the blur passes of a shadow run on a render target of a few hundred tiles, and
    // a blurred pixel costs thousands of instructions): fewer tiles per CTA, so that the pass covers every SM instead of
    // n_fb / 16 of them
    uint32_t tpc = CT_TILES;
    const uint32_t wave = (uint32_t)sm_count() * CT_MIN_CTAS_TEX * 2;
    if (n_fb < wave * CT_TILES) tpc = (n_fb + wave - 1) / wave;
    // a batch that paints with a blur filter: its blurred tiles sit side by side (the shadow's rectangle) -- one tile per
    // CTA, so that the hardware spreads them over all SMs (each is split over the CTA's four warps)
    // Textured tiles cluster (a gradient is one path), and a group of 16 of them keeps one CTA busy long after the rest of
    // the grid is done: on features.svg @ 2048^2 the slowest SM was active 58 K cycles, the average one 26 K. Up to 32 K
    // tiles -- where groups of 16 are fewer than 3.5 waves of CTAs -- groups of 8: that frame 72.3 -> 61.8 us alone,
    // 24.8 -> 25.1 us streamed (groups of 4: 65.4 / 31.4 us; 16, ordered by cost: 73.2 / 24.4 us).
    if (n_fb <= 32768u && tpc > 8u) tpc = 8u;
    if (heavy_paints) tpc = 1;
    if (tpc < 1) tpc = 1;
    ordered = whole && tpc == (uint32_t)GROUP_TILES;
    if (b.fused_fill)
        return launch_pdl(k_composite<false, true>, (n_fb + tpc - 1) / tpc, CT_THREADS, 0, s, CompositeArgs(b), p, t, clear, cc, origin, tpc, sub_tw, n_fb, ordered);
    return launch_pdl(k_composite<false, false>, (n_fb + tpc - 1) / tpc, CT_THREADS, 0, s, CompositeArgs(b), p, t, clear, cc, origin, tpc, sub_tw, n_fb, ordered);
}

}  // namespace pfcu
