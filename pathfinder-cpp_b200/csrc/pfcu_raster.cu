// fill + tile ("composite"): the two rasterizing kernels, the HBM-bound end of the path.
//
//   fill : pathfinder/shaders/d3d11/fill.comp:51-154. One warp per alpha tile; a lane owns one pixel column of two
//          4-row groups (the LUT's four channels are four consecutive rows). The tile's fills are contiguous
//          (CSR from the scatter), read once with coalesced 8-byte loads and broadcast by shuffle; coverage is
//          accumulated in registers; the 16 x 16 mask is written once, 1 byte per pixel.
//   tile : pathfinder/shaders/d3d11/tile.comp:737-850 with the shading functions of tile.comp:126-134 (combine),
//          :319-347 (radial gradient), :354-392 (blur), :459-582 (composite), :586-607 (mask), :694-726 (paint
//          metadata, decoded once per upload into a float table). One WARP per framebuffer tile (no block barriers):
//          one 16-byte load gives the tile's list range and z; the list is sorted by paint order and z-culled on
//          chip (sort.comp:49-83); a lane blends 8 pixels of one row in fp32 registers and stores them with two
//          16-byte stores. Scenes whose paints are all solid take a specialised instantiation.
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

// computeCoverage, fill.comp:51-71: coverage of the 4 rows (fragy .. fragy + 3) in one pixel column.
__device__ __forceinline__ float4 compute_coverage(const PaintView &p, float fx, float fy, float tx, float ty) {
    const bool from_left = fx < tx;
    const float lx = from_left ? fx : tx, ly = from_left ? fy : ty;
    const float rx = from_left ? tx : fx, ry = from_left ? ty : fy;
    const float wx = clampf(fx, -0.5f, 0.5f), wy = clampf(tx, -0.5f, 0.5f);
    const float offset = mixf(wx, wy, 0.5f) - lx;
    const float t = offset / (rx - lx);
    const float y = mixf(ly, ry, t);
    const float d = (ry - ly) / (rx - lx);
    const float dX = wx - wy;
    const float4 s = sample_rgba8(p.area_lut, p.lut_w, p.lut_h, (y + 8.0f) / 16.0f, fabsf(d * dX) / 16.0f, false,
                                  false, false);
    return make_float4(s.x * dX, s.y * dX, s.z * dX, s.w * dX);
}

__device__ __forceinline__ float apply_fill_rule(float cv, bool winding) {  // fill.comp:133-140
    if (winding) return clampf(fabsf(cv), 0.0f, 1.0f);
    return clampf(1.0f - fabsf(1.0f - glsl_mod(cv, 2.0f)), 0.0f, 1.0f);
}

__global__ void __launch_bounds__(256) k_fill(BatchView b, PaintView p) {
    const unsigned lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t first_alpha = b.counters->first_alpha;
    uint32_t n_alpha = *b.frame_alpha_counter - first_alpha;
    if (n_alpha > b.alpha_capacity) n_alpha = b.alpha_capacity;
    const int lx = (int)(lane & 15), g0 = (int)(lane >> 4);
    const float fragx = (float)lx + 0.5f;
    for (uint32_t a = warp; a < n_alpha; a += n_warps) {
        const uint32_t id = first_alpha + a;
        if (id >= b.mask_capacity) break;
        AlphaTile at;
        *reinterpret_cast<uint4 *>(&at) = *reinterpret_cast<const uint4 *>(&b.alpha_tiles[a]);
        const uint32_t ti = at.tile_index;
        if (ti >= b.tile_count) continue;
        const uint32_t count = at.fill_count;
        uint32_t end = b.fill_cursor[ti];
        if (end > b.fill_capacity) end = b.fill_capacity;
        const uint32_t begin = end >= count ? end - count : 0u;
        const float backdrop = (float)(int8_t)(at.packed & 0xffu);
        const bool winding = (at.packed & 0x100u) != 0;
        float4 cov0 = make_float4(backdrop, backdrop, backdrop, backdrop), cov1 = cov0;
        const float fragy0 = (float)(g0 * 4) + 0.5f, fragy1 = (float)((g0 + 2) * 4) + 0.5f;
        for (uint32_t c = begin; c < end; c += 32) {
            uint2 mine = make_uint2(0, 0);
            if (c + lane < end) mine = b.fills[c + lane];
            const int m = (int)min(32u, end - c);
            for (int k = 0; k < m; k++) {
                const uint32_t f = __shfl_sync(0xffffffffu, mine.x, k), t = __shfl_sync(0xffffffffu, mine.y, k);
                // vec4(from.x, from.y, to.x, to.y) / 256.0 - tileFragCoord.xyxy (fill.comp:87-90)
                const float fx = (float)(f & 0xffffu) / 256.0f - fragx, tx = (float)(t & 0xffffu) / 256.0f - fragx;
                const float fyq = (float)(f >> 16) / 256.0f, tyq = (float)(t >> 16) / 256.0f;
                const float4 c0 = compute_coverage(p, fx, fyq - fragy0, tx, tyq - fragy0);
                const float4 c1 = compute_coverage(p, fx, fyq - fragy1, tx, tyq - fragy1);
                cov0.x += c0.x; cov0.y += c0.y; cov0.z += c0.z; cov0.w += c0.w;
                cov1.x += c1.x; cov1.y += c1.y; cov1.z += c1.z; cov1.w += c1.w;
            }
        }
        float cv[8] = {cov0.x, cov0.y, cov0.z, cov0.w, cov1.x, cov1.y, cov1.z, cov1.w};
        uint8_t *mask = b.masks + (size_t)id * 256;
        const uint8_t *clip = at.clip_alpha >= 0 && (uint32_t)at.clip_alpha < b.mask_capacity
                                  ? b.masks + (size_t)at.clip_alpha * 256 : nullptr;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            const int row = (q < 4 ? g0 * 4 : (g0 + 2) * 4) + (q & 3);
            float v = apply_fill_rule(cv[q], winding);
            if (clip) v = fminf(v, (float)clip[row * 16 + lx] * (1.0f / 255.0f));  // fill.comp:147-150
            mask[row * 16 + lx] = (uint8_t)__float2int_rn(v * 255.0f);
        }
    }
}

cudaError_t launch_fill(const BatchView &b, const PaintView &p, cudaStream_t s) {
    if (!b.tile_count || !p.area_lut) return cudaSuccess;
    k_fill<<<sm_count() * 16, 256, 0, s>>>(b, p);
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

constexpr int MAX_SORTED = 64;      // list entries sorted in shared memory per warp; longer lists use selection
constexpr int COMPOSITE_WARPS = 4;  // framebuffer tiles per CTA

__device__ __forceinline__ uint32_t pack_rgba8(float4 c) {
    const uint32_t r = (uint32_t)__float2int_rn(clampf(c.x, 0.f, 1.f) * 255.0f);
    const uint32_t g = (uint32_t)__float2int_rn(clampf(c.y, 0.f, 1.f) * 255.0f);
    const uint32_t bl = (uint32_t)__float2int_rn(clampf(c.z, 0.f, 1.f) * 255.0f);
    const uint32_t a = (uint32_t)__float2int_rn(clampf(c.w, 0.f, 1.f) * 255.0f);
    return r | (g << 8) | (bl << 16) | (a << 24);
}

__device__ __forceinline__ float4 unpack_rgba8(uint32_t v) {
    const float k = 1.0f / 255.0f;
    return make_float4((float)(v & 0xffu) * k, (float)((v >> 8) & 0xffu) * k, (float)((v >> 16) & 0xffu) * k,
                       (float)(v >> 24) * k);
}

template <bool SOLID>
__global__ void __launch_bounds__(COMPOSITE_WARPS * 32) k_composite(BatchView b, PaintView p, TargetView tg, int clear,
                                                                    float4 clear_color) {
    __shared__ uint4 s_prims[COMPOSITE_WARPS][MAX_SORTED];
    const unsigned lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const uint32_t n_fb = (uint32_t)(b.fb_tw * b.fb_th);
    const uint32_t map = blockIdx.x * COMPOSITE_WARPS + wib;
    if (map >= n_fb) return;
    const int tile_x = (int)(map % (uint32_t)b.fb_tw), tile_y = (int)(map / (uint32_t)b.fb_tw);
    const uint4 fbt = __ldg(reinterpret_cast<const uint4 *>(&b.fb[map]));  // begin, count, z
    uint32_t n = fbt.y;
    if (n == 0 && !clear) return;  // tile.comp:743-744
    const uint32_t begin = fbt.x;
    if (begin + n > b.prim_capacity) n = begin < b.prim_capacity ? b.prim_capacity - begin : 0u;
    const int z = (int)fbt.z;

    const int row = (int)(lane >> 1), x0 = (int)(lane & 1) * 8;
    const int gx0 = tile_x * TILE + x0, gy = tile_y * TILE + row;
    const bool row_ok = gy < tg.height;
    uint32_t *dst = reinterpret_cast<uint32_t *>(tg.pixels + (size_t)gy * tg.pitch) + gx0;
    float4 dest[8];
    if (clear) {
#pragma unroll
        for (int k = 0; k < 8; k++) dest[k] = clear_color;
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++)
            dest[k] = (row_ok && gx0 + k < tg.width) ? unpack_rgba8(dst[k]) : make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // Sort by paint order and z-cull on chip (sort.comp:49-83). Keys (dense tile indices) are unique.
    uint32_t n_sorted = 0;
    const bool in_smem = n <= MAX_SORTED;
    if (in_smem && n) {
        uint4 e0 = make_uint4(0xffffffffu, 0, 0, 0), e1 = e0;
        if (lane < n) e0 = __ldg(reinterpret_cast<const uint4 *>(&b.prims[begin + lane]));
        if (lane + 32 < n) e1 = __ldg(reinterpret_cast<const uint4 *>(&b.prims[begin + lane + 32]));
        const bool keep0 = lane < n && (int)e0.x >= z, keep1 = lane + 32 < n && (int)e1.x >= z;
        uint32_t r0 = 0, r1 = 0;
        const int rounds = n > 32 ? 64 : 32;
        for (int j = 0; j < rounds; j++) {
            const uint32_t kj = __shfl_sync(0xffffffffu, j < 32 ? e0.x : e1.x, j & 31);
            const bool vj = (uint32_t)j < n && (int)kj >= z;
            r0 += (vj && kj < e0.x) ? 1u : 0u;
            r1 += (vj && kj < e1.x) ? 1u : 0u;
        }
        if (keep0) s_prims[wib][r0] = e0;
        if (keep1) s_prims[wib][r1] = e1;
        n_sorted = (uint32_t)(__popc(__ballot_sync(0xffffffffu, keep0)) + __popc(__ballot_sync(0xffffffffu, keep1)));
        __syncwarp();
    }

    ColorSampler cs;
    cs.px = p.color_px;
    cs.w = p.color_w;
    cs.h = p.color_h;
    cs.repeat_u = (p.sampling_flags & 1u) != 0;
    cs.repeat_v = (p.sampling_flags & 2u) != 0;
    cs.nearest = (p.sampling_flags & 0xcu) != 0;
    const float fragy = (float)gy + 0.5f;

    uint32_t last_key = 0;
    bool first_iter = true;
    for (uint32_t layer = 0;; layer++) {
        uint4 prim;
        if (in_smem) {
            if (layer >= n_sorted) break;
            prim = s_prims[wib][layer];
        } else {
            // selection: next smallest key >= z that is greater than the last one processed
            uint32_t best = 0xffffffffu, best_i = 0;
            for (uint32_t i = lane; i < n; i += 32) {
                const uint32_t key = __ldg(&b.prims[begin + i].key);
                if ((int)key >= z && (first_iter || key > last_key) && key < best) {
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
            if (best == 0xffffffffu) break;
            prim = __ldg(reinterpret_cast<const uint4 *>(&b.prims[begin + best_i]));
            last_key = best;
            first_iter = false;
        }
        // tile.comp:765-800
        const uint32_t color_entry = prim.z & 0xffffu;
        int tile_ctrl = (int)((prim.z >> 16) & 0xffu);
        const int backdrop = (int)prim.z >> 24;
        const int alpha = (int)prim.y;
        const uint8_t *mask = nullptr;
        if (alpha >= 0) {
            if ((uint32_t)alpha < b.mask_capacity) mask = b.masks + (size_t)alpha * 256;
        } else {
            if (backdrop != 0 && (tile_ctrl & 0x2) && (abs(backdrop) & 1) == 0) continue;  // tile.comp:786-792
            tile_ctrl &= ~0x3;
        }
        const int mask_ctrl = tile_ctrl & 0x3;
        uint2 mask8 = make_uint2(0xffffffffu, 0xffffffffu);
        if (mask_ctrl != 0 && mask) mask8 = __ldg(reinterpret_cast<const uint2 *>(mask + row * 16 + x0));
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
#pragma unroll
        for (int k = 0; k < 8; k++) {
            float mask_alpha = 1.0f;
            if (mask_ctrl != 0) {  // sampleMask, tile.comp:586-607 (backdrop is 0 for alpha tiles)
                const uint32_t m = k < 4 ? mask8.x : mask8.y;
                float cov = (float)((m >> (8 * (k & 3))) & 0xffu) * (1.0f / 255.0f);
                if (mask_ctrl & 0x1) cov = fabsf(cov);
                else cov = 1.0f - fabsf(1.0f - glsl_mod(cov, 2.0f));
                mask_alpha = fminf(mask_alpha, cov);
            }
            const float4 src = shade<SOLID>(pc, cs, (float)(gx0 + k) + 0.5f, fragy, mask_alpha, (float)tg.width,
                                            (float)tg.height);
            const float ia = 1.0f - src.w;  // tile.comp:841
            dest[k].x = dest[k].x * ia + src.x;
            dest[k].y = dest[k].y * ia + src.y;
            dest[k].z = dest[k].z * ia + src.z;
            dest[k].w = dest[k].w * ia + src.w;
        }
    }
    if (!row_ok) return;
    if (gx0 + 7 < tg.width) {
        reinterpret_cast<uint4 *>(dst)[0] =
            make_uint4(pack_rgba8(dest[0]), pack_rgba8(dest[1]), pack_rgba8(dest[2]), pack_rgba8(dest[3]));
        reinterpret_cast<uint4 *>(dst)[1] =
            make_uint4(pack_rgba8(dest[4]), pack_rgba8(dest[5]), pack_rgba8(dest[6]), pack_rgba8(dest[7]));
    } else {
#pragma unroll
        for (int k = 0; k < 8; k++)
            if (gx0 + k < tg.width) dst[k] = pack_rgba8(dest[k]);
    }
}

cudaError_t launch_composite(const BatchView &b, const PaintView &p, const TargetView &t, int clear,
                             const float clear_color[4], cudaStream_t s) {
    if (b.fb_tw <= 0 || b.fb_th <= 0) return cudaSuccess;
    const uint32_t n_fb = (uint32_t)(b.fb_tw * b.fb_th);
    const unsigned grid = (n_fb + COMPOSITE_WARPS - 1) / COMPOSITE_WARPS;
    const float4 cc = make_float4(clear_color[0], clear_color[1], clear_color[2], clear_color[3]);
    if (p.all_solid) k_composite<true><<<grid, COMPOSITE_WARPS * 32, 0, s>>>(b, p, t, clear, cc);
    else k_composite<false><<<grid, COMPOSITE_WARPS * 32, 0, s>>>(b, p, t, clear, cc);
    return cudaGetLastError();
}

}  // namespace pfcu
