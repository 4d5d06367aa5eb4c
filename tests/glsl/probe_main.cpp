// Driver for tests/test_glsl_shim.py: runs the probe shader over its 4 x 2 invocations and prints the image bytes.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "glsl_shim.h"

#include "probe_comp.inc"

int main() {
    using namespace probe_comp;
    std::vector<float> tex = {  // 3 x 2 RGBA texels
        0.10f, 0.20f, 0.30f, 1.00f, 0.90f, 0.10f, 0.50f, 0.25f, 0.40f, 0.80f, 0.00f, 0.75f,
        0.00f, 1.00f, 0.60f, 0.50f, 0.30f, 0.30f, 0.30f, 0.30f, 1.00f, 0.00f, 0.20f, 0.10f};
    const uint32_t words[8] = {0x10204080u, 0xff00ff00u, 0x01020304u, 0x7f7f7f7fu, 0xdeadbeefu, 0x00000000u, 0xffffffffu, 0x80402010u};
    uint8_t out[8 * 4] = {0};
    uint32_t counters[9] = {0};
    int32_t isigned[2] = {-5, 9};
    for (int variant = 0; variant < 2; variant++) {
        uTex = glsl::sampler2D{tex.data(), 3, 2, variant == 0, variant == 1, false};
        uOut.texels = out;
        uOut.width = 4;
        uOut.height = 2;
        uScale = glsl::vec4(1.0f, 0.5f, 2.0f, 0.75f);
        uOrigin = glsl::ivec2(0, 0);
        iIn = words;
        iCounters = counters;
        iSigned = isigned;
        for (unsigned y = 0; y < (unsigned)LOCAL_Y; y++)
            for (unsigned x = 0; x < (unsigned)LOCAL_X; x++) {
                gl_LocalInvocationID.x = x;
                gl_LocalInvocationID.y = y;
                shader_main();
            }
        for (int i = 0; i < 32; i++) printf("%d%c", out[i], i == 31 ? '\n' : ' ');
    }
    for (int i = 0; i < 9; i++) printf("%u%c", counters[i], i == 8 ? '\n' : ' ');
    printf("%d %d\n", isigned[0], isigned[1]);
    // sub-texel precision: a LINEAR fetch ~1e-6 texel off a texel centre returns that texel exactly
    uTex = glsl::sampler2D{tex.data(), 3, 2, true, false, false};
    const glsl::vec4 c = glsl::texture(uTex, glsl::vec2((1.0f + 0.5f) * (1.0f / 3.0f) + 1e-7f, 0.25f));
    printf("%.9g %.9g %.9g %.9g\n", c.x, c.y, c.z, c.w);
    return 0;
}
