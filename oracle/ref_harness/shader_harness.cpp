// TEST INFRASTRUCTURE (oracle side). Runs the reference's own compute shaders on the CPU.
//
// oracle/Makefile generates <name>_comp.inc from pathfinder/shaders/d3d11/<name>.comp where they lie under
// /root/reference (make_shader_cpp.py: a syntactic GLSL -> C++ rewrite over glsl_shim.h) into a temporary directory and
// compiles this file into oracle/_ref/libpfshader.so. The entry points bind the shaders' buffers / textures / uniforms to host
// memory exactly as RendererD3D11 binds them (core/d3d11/renderer.cpp:365-448 draw_tiles, :959-1002 draw_fills) and
// run every work group, one invocation after the other. What comes out is the reference's GPU-driven pixel pipeline
// evaluated in IEEE fp32 -- the thing tests/ pins oracle/pf_oracle.c's fill / tile restatement against.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "glsl_shim.h"

#include "fill_comp.inc"
#include "tile_comp.inc"
#include "propagate_comp.inc"
#include "sort_comp.inc"
#include "dice_comp.inc"
#include "bound_comp.inc"
#include "bin_comp.inc"

namespace {

// RGBA8 -> the float texels a sampler returns (unorm: c / 255)
std::vector<float> unorm_texels(const uint8_t *rgba, size_t n_texels) {
    std::vector<float> out(n_texels * 4);
    for (size_t i = 0; i < n_texels * 4; i++) out[i] = (float)rgba[i] / 255.0f;
    return out;
}

}  // namespace

extern "C" {

int pfshader_abi_version(void) { return 1; }

/// Diagnostic: compute every bilinear weight with `bits` fraction bits (8 = what GPUs do); 0 = exact fp32 (default).
void pfshader_set_subtexel_bits(int bits) { glsl::subtexel_bits() = bits; }

/// fill.comp over the alpha tiles [first_alpha, first_alpha + n_alpha) of one batch (draw_fills, renderer.cpp:959-1002:
/// one work group of 16 x 4 invocations per alpha tile; uAlphaTileRange = the batch's range of frame-global indices).
///   fills       : 3 x u32 per fill {from, to, next fill index or -1} (bin.comp's linked lists)
///   tiles       : 4 x u32 per dense tile (TileD3D11, gpu_data.h:38-47)
///   alpha_tiles : 2 x u32 per alpha tile of the batch {dense tile index, clip alpha tile index or -1}
///   area_lut    : 256 x 256 RGBA8;  mask: RGBA8 image of mask_w x mask_h texels (read for clips, written)
int pfshader_fill(const uint32_t *fills, const uint32_t *tiles, const uint32_t *alpha_tiles, int first_alpha, int n_alpha,
                  const uint8_t *area_lut, int lut_w, int lut_h, uint8_t *mask, int mask_w, int mask_h) {
    using namespace fill_comp;
    const std::vector<float> lut = unorm_texels(area_lut, (size_t)lut_w * lut_h);
    iFills = fills;
    iTiles = tiles;
    iAlphaTiles = alpha_tiles;
    uAlphaTileRange = glsl::ivec2(first_alpha, first_alpha + n_alpha);
    uAreaLUT.texels = lut.data();
    uAreaLUT.width = lut_w;
    uAreaLUT.height = lut_h;
    uAreaLUT.linear = true;  // core/renderer.cpp:117-165: linear filter, clamp to edge
    uAreaLUT.repeat_u = uAreaLUT.repeat_v = false;
    uDest.texels = mask;
    uDest.width = mask_w;
    uDest.height = mask_h;
    for (int wg = 0; wg < n_alpha; wg++) {
        gl_WorkGroupID.x = (unsigned)wg & 0x7fffu;  // the 64 K dispatch workaround, renderer.cpp:993-996
        gl_WorkGroupID.y = (unsigned)wg >> 15;
        gl_WorkGroupID.z = 0;
        for (unsigned ly = 0; ly < (unsigned)LOCAL_Y; ly++)
            for (unsigned lx = 0; lx < (unsigned)LOCAL_X; lx++) {
                gl_LocalInvocationID.x = lx;
                gl_LocalInvocationID.y = ly;
                gl_LocalInvocationID.z = 0;
                shader_main();
            }
    }
    uAreaLUT.texels = nullptr;
    return 0;
}

/// tile.comp over every framebuffer tile (draw_tiles, renderer.cpp:365-448: one work group of 16 x 4 invocations per
/// tile of the destination's tile grid; default sampler = linear + clamp to edge, the colour texture's by its flags).
///   tiles          : 4 x u32 per dense tile, word 0 = next tile of the framebuffer tile's SORTED list (sort.comp)
///   first_tile_map : head of every framebuffer tile's list or -1
///   metadata       : RGBA fp32 texels of the RGBA16F paint metadata texture (md_w x md_h)
///   color          : RGBA8 colour texture (pattern page) or a 1 x 1 dummy; sampling_flags: TextureSamplingFlags
///   mask           : the RGBA8 mask texture fill.comp wrote;  dest: RGBA8 image (destination or render-target page)
int pfshader_tile(const uint32_t *tiles, const int32_t *first_tile_map, int fb_tw, int fb_th, const float *metadata,
                  int md_w, int md_h, const uint8_t *color, int color_w, int color_h, uint32_t sampling_flags,
                  const uint8_t *mask, int mask_w, int mask_h, uint8_t *dest, int dest_w, int dest_h, int load_action,
                  const float *clear_color) {
    using namespace tile_comp;
    // PFSHADER_DEBUG="tile_x,tile_y,local_x,local_y": trace every texture() call of one invocation to stderr
    int dbg_v[4] = {0, 0, 0, 0};
    const char *dbg_env = getenv("PFSHADER_DEBUG");
    const bool dbg = dbg_env && sscanf(dbg_env, "%d,%d,%d,%d", &dbg_v[0], &dbg_v[1], &dbg_v[2], &dbg_v[3]) == 4;
    const std::vector<float> color_f = unorm_texels(color, (size_t)color_w * color_h);
    const std::vector<float> mask_f = unorm_texels(mask, (size_t)mask_w * mask_h);
    const float zero[4] = {0.f, 0.f, 0.f, 0.f};
    iTiles = tiles;
    iFirstTileMap = first_tile_map;
    uTextureMetadata = glsl::sampler2D{metadata, md_w, md_h, true, false, false};
    uZBuffer = glsl::sampler2D{zero, 1, 1, true, false, false};
    uGammaLUT = glsl::sampler2D{zero, 1, 1, true, false, false};
    uColorTexture0 = glsl::sampler2D{color_f.data(), color_w, color_h, (sampling_flags & 0xcu) == 0,
                                     (sampling_flags & 1u) != 0, (sampling_flags & 2u) != 0};
    uMaskTexture0 = glsl::sampler2D{mask_f.data(), mask_w, mask_h, true, false, false};
    uDestImage.texels = dest;
    uDestImage.width = dest_w;
    uDestImage.height = dest_h;
    uClearColor = glsl::vec4(clear_color[0], clear_color[1], clear_color[2], clear_color[3]);
    uLoadAction = load_action;
    uTileSize = glsl::vec2(16.0f, 16.0f);
    uTextureMetadataSize = glsl::vec2((float)md_w, (float)md_h);
    uFramebufferSize = glsl::vec2((float)dest_w, (float)dest_h);
    uFramebufferTileSize = glsl::ivec2(fb_tw, fb_th);
    uMaskTextureSize0 = glsl::vec2((float)mask_w, (float)mask_h);
    uColorTextureSize0 = glsl::vec2((float)color_w, (float)color_h);
    for (int ty = 0; ty < fb_th; ty++)
        for (int tx = 0; tx < fb_tw; tx++) {
            gl_WorkGroupID.x = (unsigned)tx;
            gl_WorkGroupID.y = (unsigned)ty;
            gl_WorkGroupID.z = 0;
            for (unsigned ly = 0; ly < (unsigned)LOCAL_Y; ly++)
                for (unsigned lx = 0; lx < (unsigned)LOCAL_X; lx++) {
                    gl_LocalInvocationID.x = lx;
                    gl_LocalInvocationID.y = ly;
                    gl_LocalInvocationID.z = 0;
                    glsl::debug_trace() = dbg && tx == dbg_v[0] && ty == dbg_v[1] && (int)lx == dbg_v[2] && (int)ly == dbg_v[3];
                    shader_main();
                }
        }
    glsl::debug_trace() = false;
    uColorTexture0.texels = uMaskTexture0.texels = uTextureMetadata.texels = nullptr;
    return 0;
}

/// propagate.comp over every tile column of one batch (propagate_tiles, renderer.cpp:853-953: one invocation per column).
///   draw_metadata : PropagateMetadataD3D11 records (3 x uvec4 per path);  backdrops: BackdropInfoD3D11 (3 x i32 per
///                   column) with the counts bin left in [0];  draw_tiles: 4 x u32 per dense tile, read and written
///   clip_metadata / clip_tiles: the clip batch's records or NULL. NOTE: the shader indexes clip_metadata with a stride of
///                   TWO uvec4 (propagate.comp:110-111) while the reference binds the clip batch's 48-byte records
///                   (renderer.cpp:910-916), so only clip path 0 is read correctly upstream; the caller passes what the
///                   reference binds.
///   z_buffer      : 8 + framebuffer tiles i32 ([4] = alpha tile counter), first_tile_map: framebuffer tiles i32 (-1),
///   alpha_tiles   : 2 x u32 per alpha tile (written).
int pfshader_propagate(const uint32_t *draw_metadata, const uint32_t *clip_metadata, const int32_t *backdrops,
                       uint32_t *draw_tiles, uint32_t *clip_tiles, int32_t *z_buffer, int32_t *first_tile_map,
                       uint32_t *alpha_tiles, int fb_tw, int fb_th, int column_count, int first_alpha) {
    using namespace propagate_comp;
    iDrawMetadata = reinterpret_cast<const glsl::uvec4 *>(draw_metadata);
    iClipMetadata = reinterpret_cast<const glsl::uvec4 *>(clip_metadata);
    iBackdrops = backdrops;
    iDrawTiles = draw_tiles;
    iClipTiles = clip_tiles;
    iZBuffer = z_buffer;
    iFirstTileMap = first_tile_map;
    iAlphaTiles = alpha_tiles;
    uFramebufferTileSize = glsl::ivec2(fb_tw, fb_th);
    uColumnCount = column_count;
    uFirstAlphaTileIndex = first_alpha;
    for (int c = 0; c < column_count; c++) {
        gl_GlobalInvocationID.x = (unsigned)c;
        gl_GlobalInvocationID.y = gl_GlobalInvocationID.z = 0;
        shader_main();
    }
    return 0;
}

/// sort.comp over every framebuffer tile (sort_tiles, renderer.cpp:1004-1039).
int pfshader_sort(uint32_t *tiles, int32_t *first_tile_map, const int32_t *z_buffer, int fb_tiles) {
    using namespace sort_comp;
    iTiles = tiles;
    iFirstTileMap = first_tile_map;
    iZBuffer = z_buffer;
    uTileCount = fb_tiles;
    for (int t = 0; t < fb_tiles; t++) {
        gl_GlobalInvocationID.x = (unsigned)t;
        gl_GlobalInvocationID.y = gl_GlobalInvocationID.z = 0;
        shader_main();
    }
    return 0;
}

/// dice.comp (dice_segments, renderer.cpp:618-731): one invocation per batch segment. transform = m11 m21 m12 m22 m13 m23.
/// indirect[3] counts the microlines (may exceed max_microlines: the caller retries with more room, as the reference does).
int pfshader_dice(uint32_t *indirect, const uint32_t *dice_metadata, const float *points, int point_count,
                  const uint32_t *indices, uint32_t *microlines, const float *transform, int path_count, int segment_count,
                  int max_microlines) {
    using namespace dice_comp;
    iComputeIndirectParams = indirect;
    iDiceMetadata = reinterpret_cast<const glsl::uvec4 *>(dice_metadata);
    // (glsl::vec2 carries its swizzle proxies and is wider than two floats: repack the std430 array)
    std::vector<glsl::vec2> point_array;
    {
        point_array.resize((size_t)point_count);
        for (size_t i = 0; i < (size_t)point_count; i++) point_array[i] = glsl::vec2(points[i * 2], points[i * 2 + 1]);
    }
    iPoints = point_array.data();
    iInputIndices = reinterpret_cast<const glsl::uvec2 *>(indices);
    iMicrolines = reinterpret_cast<glsl::uvec4 *>(microlines);
    uTransform = glsl::mat2(transform[0], transform[1], transform[2], transform[3]);
    uTranslation = glsl::vec2(transform[4], transform[5]);
    uPathCount = path_count;
    uLastBatchSegmentIndex = segment_count;
    uMaxMicrolineCount = max_microlines;
    for (int i = 0; i < segment_count; i++) {
        gl_GlobalInvocationID.x = (unsigned)i;
        shader_main();
    }
    return 0;
}

/// bound.comp (bound, renderer.cpp:733-775): one invocation per dense tile; tile_path_info = TilePathInfoD3D11 records.
int pfshader_bound(const uint32_t *tile_path_info, uint32_t *tiles, int path_count, int tile_count) {
    using namespace bound_comp;
    iTilePathInfo = reinterpret_cast<const glsl::uvec4 *>(tile_path_info);
    iTiles = tiles;
    uPathCount = path_count;
    uTileCount = tile_count;
    for (int i = 0; i < tile_count; i++) {
        gl_GlobalInvocationID.x = (unsigned)i;
        shader_main();
    }
    return 0;
}

/// bin.comp (bin_segments, renderer.cpp:777-851): one invocation per microline. indirect_draw = the z-buffer's 8-word
/// header ([1] counts the fills; may exceed max_fills: retry with more room). backdrops: 3 x u32 per column, [0] updated.
int pfshader_bin(const uint32_t *microlines, const int32_t *metadata, uint32_t *indirect_draw, uint32_t *fills,
                 uint32_t *tiles, uint32_t *backdrops, int microline_count, int max_fills) {
    using namespace bin_comp;
    iMicrolines = reinterpret_cast<const glsl::uvec4 *>(microlines);
    iMetadata = reinterpret_cast<const glsl::ivec4 *>(metadata);
    iIndirectDrawParams = indirect_draw;
    iFills = fills;
    iTiles = tiles;
    iBackdrops = backdrops;
    uMicrolineCount = microline_count;
    uMaxFillCount = max_fills;
    for (int i = 0; i < microline_count; i++) {
        gl_GlobalInvocationID.x = (unsigned)i;
        shader_main();
    }
    return 0;
}

}  // extern "C"
