// Application-side host loop over the C-ABI (include/pfcu.h): a sequence of frames of one scene rendered with several frames
// in flight, one renderer context per frame in flight -- what an application that animates or batch-renders does on top of
// pfcu_submit_frame / pfcu_wait_frame, written in C++ like the reference's own applications (demo/native/main.cpp drives
// Canvas::draw from a C++ loop). Every frame uploads the scene's segments from the caller's host buffers
// (RendererD3D11::upload_scene, d3d11/renderer.cpp:314), prepares the clip batches in reverse and every draw batch, draws, and
// reads its counters back, exactly the sequence of RendererD3D11::draw (d3d11/renderer.cpp:302-336).
//
// bench.py's e2e leg and tools/e2e_stream.py call it through ctypes; the same loop in Python (pfcu.py Renderer.draw) costs
// 46 us of host time per frame, more than the GPU needs for most of a tiger.svg frame.
#pragma once

#include <stdint.h>

#include "../../include/pfcu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pfhost_draw {
    uint32_t batch_id;
    int32_t target_page;      // < 0: the destination
    int32_t color_page;       // < 0: none
    uint32_t sampling_flags;
} pfhost_draw;

typedef struct pfhost_scene {
    const float *points[2];        // draw, clip
    const uint32_t *indices[2];
    uint32_t n_points[2], n_segments[2];
    const pfcu_batch_desc *clip_batches;  // in scene order; prepared in reverse, empty ones skipped
    uint32_t n_clip_batches;
    const pfcu_batch_desc *draw_batches;
    const pfhost_draw *draws;             // one per draw batch
    uint32_t n_draw_batches;
    float clear_color[4];
} pfhost_scene;

/* Renders n_frames frames of `scene`, frame i on contexts[i % n_contexts], waiting for a context's previous frame before it
 * reuses it. pixels (may be NULL): n_contexts page-locked buffers; when given, every frame's target is also read back
 * (pfcu_read_target_async) into its context's buffer. Returns PFCU_OK or the first error; wall_seconds = the whole loop
 * including the final waits; last = statistics of the last frame; retries = replays summed over all frames. */
int pfhost_stream_frames(pfcu_ctx *const *contexts, uint32_t n_contexts, const pfhost_scene *scene, uint32_t n_frames,
                         uint8_t *const *pixels, double *wall_seconds, pfcu_frame_stats *last, uint32_t *retries);

#ifdef __cplusplus
}
#endif
